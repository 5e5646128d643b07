"""bench/gpu_sat.py -- saturation experiment: many async solve launches over S streams, lean host loop."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mpc_ros_b200 import capi
from bench import gen_py

def main():
    B = int(sys.argv[1]); S = int(sys.argv[2]); K = int(sys.argv[3]); maxc = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    L = capi.lib(); dev = torch.device("cuda:0")
    prm = capi.yaml_default_params(); prm.delay_mode = 0
    if len(sys.argv) > 11: prm.max_iter = int(sys.argv[11])
    sv = capi.Solver(prm, B, 0)
    if hasattr(L, 'mpc_b200_debug_profile'): L.mpc_b200_debug_profile(sv._h, None)
    if maxc: sv.set_option("max_ctas", maxc)
    pbo = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    if pbo: sv.set_option("problems_per_cta", pbo)
    if len(sys.argv) > 6: sv.set_option("hard_first", int(sys.argv[6]))
    if len(sys.argv) > 7: sv.set_option("dual_groups", int(sys.argv[7]))
    R = int(sys.argv[9]) if len(sys.argv) > 9 else 8       # distinct input / output sets the launches rotate over (48: larger than L2, as bench.py)
    g = gen_py.problems(20261020, B * R)
    M = g["M"]
    def split(x): return [torch.from_numpy(np.ascontiguousarray(x[:, j*B:(j+1)*B])).to(dev) for j in range(R)]
    wx, wy, pose, vel = split(g["wx"]), split(g["wy"]), split(g["pose"]), split(g["vel"])
    f64 = dict(dtype=torch.float64, device=dev)
    coef = [torch.zeros((4, B), **f64) for _ in range(R)]; state = [torch.zeros((6, B), **f64) for _ in range(R)]
    u0 = [torch.zeros((2, B), **f64) for _ in range(R)]; pred = [torch.zeros((60, B), **f64) for _ in range(R)]
    # one status array per LAUNCH, pre-filled with a sentinel: a launch that did not solve its batch shows up at the end
    stat_all = torch.full((K + S, B), -7, dtype=torch.int32, device=dev)
    stat = [stat_all[K + j] for j in range(R)]
    # (torch.cuda.Stream() hands out at most 32 distinct streams per device: the library creates real ones)
    raw = [capi.stream_create(0) for _ in range(S)]
    streams = [torch.cuda.ExternalStream(q, device=dev) for q in raw]
    for j in range(R):
        sv.prestep_raw(B, M, wx[j], wy[j], pose[j], vel[j], coef[j], state[j])
    torch.cuda.synchronize()
    args = []
    full_out = len(sys.argv) > 10 and int(sys.argv[10]) != 0     # also obj / iters / kkt per problem (as bench.py)
    obj = [torch.zeros(B, **f64) for _ in range(R)]; kkt = [torch.zeros(B, **f64) for _ in range(R)]
    its = [torch.zeros(B, dtype=torch.int32, device=dev) for _ in range(R)]
    for j in range(R):
        args.append((sv._h, B, state[j].data_ptr(), coef[j].data_ptr(), None, None, u0[j].data_ptr(), pred[j].data_ptr(),
                     obj[j].data_ptr() if full_out else None, stat[j].data_ptr(), its[j].data_ptr() if full_out else None,
                     kkt[j].data_ptr() if full_out else None, None))
    sp = raw
    with_pre = len(sys.argv) > 8 and int(sys.argv[8]) != 0       # also the pre-step kernel in front of every solve (as bench.py)
    pre_args = [(sv._h, B, M, wx[j].data_ptr(), wy[j].data_ptr(), pose[j].data_ptr(), vel[j].data_ptr(), coef[j].data_ptr(), state[j].data_ptr()) for j in range(R)]
    f = L.mpc_b200_solve_batch
    for j in range(S):
        f(*args[j % R], sp[j % S])
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams: s.wait_event(e0)
    t0 = time.perf_counter()
    # at most 2 launches in flight per stream (the handle refuses more launches in flight than it has ring slots)
    ev = [[torch.cuda.Event() for _ in range(2)] for _ in range(S)]
    for j in range(K):
        s_ = j % S; k_ = j // S
        if k_ >= 2: ev[s_][k_ % 2].synchronize()
        a_ = list(args[j % R]); a_[9] = stat_all[j].data_ptr()
        if with_pre: L.mpc_b200_prestep_batch(*pre_args[j % R], sp[s_])
        rc = f(*a_, sp[s_])
        if rc != 0: raise RuntimeError("solve_batch failed: %d" % rc)
        ev[s_][k_ % 2].record(streams[s_])
    t1 = time.perf_counter()
    for s in streams: torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if hasattr(L, "mpc_b200_debug_profile"):
        buf = (C.c_longlong * 1024)()
        if L.mpc_b200_debug_profile(sv._h, buf) and buf[1000] > 0:
            print("   avg global cycle: %.0f SM cycles (%d cycles total over all CTAs of all launches)" % (buf[1001] / buf[1000], buf[1000]))
            print("   ctrl_decide parts (thread 0): reduce %.0f, line search %.0f, error+tests %.0f, mu+rest %.0f" % tuple(buf[1020 + i] / buf[1000] for i in range(4)))
            if buf[1002] > 0: print("   busy lanes per cycle: %.2f" % (buf[1002] / buf[1000]))
            names = ["-", "refill", "P3_apply_coeffs", "P4_backward", "P4_forward(+rollout,prefetch)", "P5_step", "P6", "P1_eval", "P2_rest", "P6_adjoint", "P2_ctrl_decide"]
            print("   per-cycle breakdown (control thread 0 of every CTA): " + ", ".join("%s %.0f" % (names[i], buf[1008 + i] / buf[1000]) for i in range(1, 11)))
    conv = int((stat_all[:K] == 1).sum()); unwritten = int((stat_all[:K] == -7).sum())
    print("   converged: %d of %d problems over all %d timed launches; problems never written: %d" % (conv, K * B, K, unwritten))
    print("B %d S %d K %d maxctas %d: host issue %.1f us/launch, device %.3f ms/launch, %.2f M solves/s" %
          (B, S, K, maxc, (t1 - t0) / K * 1e6, ms / K, B * K / ms / 1e3))
main()
