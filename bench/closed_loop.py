"""bench/closed_loop.py -- BASELINE config 5: closed-loop simulation, R robots x T control ticks.

Per tick (the reference's control tick, SURVEY section 3A, with the ROS plumbing replaced by arrays):
  window the reference path ahead of the robot and down-sample it (mpc_planner_ros.cpp:266-291, :365-391)
  -> pre-step: transform + polyfit + delay-compensated state (driving_state.cpp:196-256)   [GPU, K1]
  -> MPC::Solve, warm-started from the previous tick's shifted solution                     [GPU, K3]
  -> speed command clamp (driving_state.cpp:263-269) -> unicycle plant step.
run_gpu drives the loop from the host (windowing and plant in numpy); run_gpu_device keeps everything on the
device: windowing (mpc_b200_window_batch), post-step (mpc_b200_poststep_batch) and the warm-start shift are
kernels too, robots are split over independent streams.

    python bench/closed_loop.py [robots] [ticks] [--cold] [--oracle-subset K] [--device]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from bench import gen_py  # noqa: E402

DS = 0.05


class Fleet:
    """R robots on the three synthetic tracks (robot i on track i % 3)."""

    def __init__(self, R, seed=20261018 + 5, path_length=5.0):
        self.R = R
        self.paths = [gen_py.path(k) for k in range(3)]
        rng = np.random.default_rng(seed)
        self.kind = np.arange(R) % 3
        self.idx = np.zeros(R, dtype=np.int64)
        self.pose = np.zeros((3, R))
        self.v = np.zeros(R); self.w = np.zeros(R); self.thr = np.zeros(R)
        for i in range(R):
            px, py = self.paths[self.kind[i]]
            n = len(px)
            j = int(rng.integers(n))
            tx, ty = px[(j + 1) % n] - px[j], py[(j + 1) % n] - py[j]
            tl = np.hypot(tx, ty)
            lat = rng.uniform(-0.2, 0.2)
            self.idx[i] = j
            self.pose[:, i] = [px[j] - ty / tl * lat, py[j] + tx / tl * lat, np.arctan2(ty, tx) + rng.uniform(-0.3, 0.3)]
            self.v[i] = rng.uniform(0.0, 0.4)
        self.win = int(round(path_length / DS))
        self.step = int(path_length / 10.0 / DS)
        self.M = gen_py._lib().mpcgen_num_waypoints(path_length)

    def window(self):
        """Cut the plan and down-sample it (mpc_planner_ros.cpp:266-291, :365-391), host side of run_gpu.  The rule is
        the reference's: plan points are erased from the front while the squared distance to the robot does not grow
        (start value 10e5), so the plan begins at the first point that is farther away than its predecessor; at most 64
        points per tick (the kernel's cap)."""
        R, M = self.R, self.M
        wx = np.zeros((M, R)); wy = np.zeros((M, R))
        for i in range(R):
            px, py = self.paths[self.kind[i]]
            n = len(px)
            cand = (self.idx[i] + np.arange(0, 65)) % n
            d2 = (px[cand] - self.pose[0, i]) ** 2 + (py[cand] - self.pose[1, i]) ** 2
            prev = np.concatenate([[10e5], d2[:-1]])
            grow = np.nonzero(prev[:64] < d2[:64])[0]
            self.idx[i] = cand[int(grow[0])] if len(grow) else cand[64]
            sel = list(range(0, self.win, self.step)) + [self.win - 1]
            q = (self.idx[i] + np.array(sel)) % n
            wx[:, i] = px[q]; wy[:, i] = py[q]
        return wx, wy

    def vel(self):
        return np.stack([self.v, self.w, self.thr])

    def track_distance(self):
        """True distance of every robot to its track (the logged cte is c[0] of the cubic fit, which is meaningless
        where a 5 m window wraps around a sharp corner)."""
        d = np.zeros(self.R)
        for i in range(self.R):
            px, py = self.paths[self.kind[i]]
            d[i] = np.sqrt(((px - self.pose[0, i]) ** 2 + (py - self.pose[1, i]) ** 2).min())
        return d

    def actuate(self, w, thr, dt, ref_v):
        """driving_state.cpp:263-269 then a unicycle plant."""
        speed = self.v + thr * dt
        speed = np.minimum(speed, ref_v)
        self.w = w.copy(); self.thr = thr.copy()
        self.pose[0] += speed * np.cos(self.pose[2]) * dt
        self.pose[1] += speed * np.sin(self.pose[2]) * dt
        self.pose[2] += w * dt
        self.pose[2] = (self.pose[2] + np.pi) % (2 * np.pi) - np.pi
        self.v = speed


def run_gpu(R, T, warm=True, seed=20261018 + 5, record_solver=False):
    import torch
    from mpc_ros_b200 import capi
    prm = capi.yaml_default_params()          # delay_mode true, as in mpc_params.yaml:4
    N = prm.mpc_steps; dt = prm.dt
    sv = capi.Solver(prm, R, 0)
    fleet = Fleet(R, seed)
    dev = torch.device("cuda:0")
    ws = capi.lib().mpc_b200_warm_size(N)
    f64 = dict(dtype=torch.float64, device=dev)
    warm_a = torch.zeros((ws, R), **f64); warm_b = torch.zeros((ws, R), **f64)
    d_state = torch.zeros((6, R), **f64); d_coef = torch.zeros((4, R), **f64)
    d_u0 = torch.zeros((2, R), **f64); d_pred = torch.zeros((3 * N, R), **f64)
    d_it = torch.zeros(R, dtype=torch.int32, device=dev); d_st = torch.zeros(R, dtype=torch.int32, device=dev)
    trace = dict(cte=[], eth=[], iters=[], conv=[], w=[], thr=[], dist=[])
    t_solve = 0.0
    for t in range(T):
        wx, wy = fleet.window()
        d_wx = torch.from_numpy(wx).to(dev); d_wy = torch.from_numpy(wy).to(dev)
        d_pose = torch.from_numpy(np.ascontiguousarray(fleet.pose)).to(dev)
        d_vel = torch.from_numpy(np.ascontiguousarray(fleet.vel())).to(dev)
        # tracking errors as the reference logs them (cte = c[0], etheta before delay compensation)
        cte_e = torch.zeros((2, R), **f64)
        sv.polyfit_raw(R, fleet.M, d_wx, d_wy, d_pose, d_coef, cte_e)
        sv.prestep_raw(R, fleet.M, d_wx, d_wy, d_pose, d_vel, d_coef, d_state)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        use_warm = warm and t > 0
        sv.solve_raw(R, d_state, d_coef, d_u0, d_pred, warm_in=warm_a if use_warm else None, status=d_st, iters=d_it,
                     warm_out=warm_b)
        torch.cuda.synchronize()
        t_solve += time.perf_counter() - t0
        sv.warm_shift(R, warm_b, warm_a)
        u0 = d_u0.cpu().numpy()
        ce = cte_e.cpu().numpy()
        trace["cte"].append(ce[0].copy()); trace["eth"].append(ce[1].copy())
        trace["iters"].append(d_it.cpu().numpy().copy()); trace["conv"].append((d_st.cpu().numpy() == 1))
        trace["w"].append(u0[0].copy()); trace["thr"].append(u0[1].copy())
        trace["dist"].append(fleet.track_distance())
        fleet.actuate(u0[0], u0[1], dt, prm.ref_vel)
    sv.close()
    out = {k: np.array(v) for k, v in trace.items()}
    out["solve_s"] = t_solve
    out["kind"] = fleet.kind.copy()
    return out


TRACKS = ("infinity", "epitrochoid", "square")


def per_track(kind, cte, dist):
    """Tracking statistics per track: |cte| as the reference logs it (c[0] of the fit) and the true distance to the track."""
    out = {}
    for k, name in enumerate(TRACKS):
        m = kind == k
        if not m.any():
            continue
        a = np.abs(cte[:, m]); d = dist[:, m]
        out[name] = dict(robots=int(m.sum()), mean_abs_cte=float(a.mean()), median_abs_cte=float(np.median(a)), max_abs_cte=float(a.max()),
                         frac_ticks_abs_cte_gt_1m=float((a > 1.0).mean()), mean_dist=float(d.mean()), median_dist=float(np.median(d)),
                         max_dist=float(d.max()), frac_ticks_dist_gt_1m=float((d > 1.0).mean()),
                         frac_robots_ever_dist_gt_1m=float((d.max(axis=0) > 1.0).mean()))
    return out


def run_gpu_device(R, T, groups=8, seed=20261018 + 5, max_iter=100, trace=True):
    """Device-resident loop: windowing, pre-step, warm-started solve, warm shift and post-step are C-ABI
    kernels, the unicycle plant (simulation, not part of the reference) included.
    Robots are split into `groups` independent streams so that a slow solve only delays its own group."""
    import torch
    from mpc_ros_b200 import capi
    prm = capi.yaml_default_params()
    prm.max_iter = max_iter
    N = prm.mpc_steps; dt = prm.dt
    dev = torch.device("cuda:0")
    fleet = Fleet(R, seed)
    M = capi.lib().mpc_b200_num_waypoints(prm)
    assert M == fleet.M
    f64 = dict(dtype=torch.float64, device=dev); i32 = dict(dtype=torch.int32, device=dev)
    lens = [len(p[0]) for p in fleet.paths]
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    d_px = torch.from_numpy(np.concatenate([p[0] for p in fleet.paths])).to(dev)
    d_py = torch.from_numpy(np.concatenate([p[1] for p in fleet.paths])).to(dev)
    d_off = torch.from_numpy(offs).to(dev); d_len = torch.tensor(lens, **i32)
    ws = capi.lib().mpc_b200_warm_size(N)
    G = max(1, min(groups, R))
    bounds = [(g * R // G, (g + 1) * R // G) for g in range(G)]
    sv = capi.Solver(prm, max(hi - lo for lo, hi in bounds), 0)
    sv.set_option("max_ctas", max(1, 148 // G))
    grp = []
    for lo, hi in bounds:
        B = hi - lo
        st = torch.cuda.Stream()
        grp.append(dict(B=B, st=st, sp=st.cuda_stream,
                        tid=torch.from_numpy(fleet.kind[lo:hi].astype(np.int32)).to(dev),
                        idx=torch.from_numpy(fleet.idx[lo:hi].astype(np.int32)).to(dev),
                        pose=torch.from_numpy(np.ascontiguousarray(fleet.pose[:, lo:hi])).to(dev),
                        vel=torch.from_numpy(np.ascontiguousarray(fleet.vel()[:, lo:hi])).to(dev),
                        wx=torch.zeros((M, B), **f64), wy=torch.zeros((M, B), **f64), coef=torch.zeros((4, B), **f64),
                        state=torch.zeros((6, B), **f64), u0=torch.zeros((2, B), **f64), pred=torch.zeros((3 * N, B), **f64),
                        cmd=torch.zeros((2, B), **f64), wa=torch.zeros((ws, B), **f64), wb=torch.zeros((ws, B), **f64),
                        it=torch.zeros(B, **i32), stt=torch.zeros(B, **i32), ce=torch.zeros((2, B), **f64),
                        cte=[], iters=[], conv=[]))
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for g in grp:
        g["st"].wait_event(e0)
    # the host runs ahead of the device, but not without bound: a handle refuses more launches in flight than its
    # queue ring has slots (1,024; here 8 launches per tick), so every 16 ticks the host marks the streams and waits
    # for the marks set 32 ticks earlier: at most 48 ticks = 384 launches are in flight
    marks = []
    for t in range(T):
        if t % 16 == 0:
            cur = []
            for g in grp:
                ev = torch.cuda.Event(); ev.record(g["st"]); cur.append(ev)
            marks.append(cur)
            while len(marks) > 2:
                for ev in marks.pop(0):
                    ev.synchronize()
        for g in grp:
            B, sp = g["B"], g["sp"]
            with torch.cuda.stream(g["st"]):
                sv.window_raw(B, d_px, d_py, d_off, d_len, g["tid"], g["idx"], g["pose"], g["wx"], g["wy"], stream=sp)
                if trace:
                    sv.polyfit_raw(B, M, g["wx"], g["wy"], g["pose"], g["coef"], g["ce"], stream=sp)
                sv.prestep_raw(B, M, g["wx"], g["wy"], g["pose"], g["vel"], g["coef"], g["state"], stream=sp)
                sv.solve_raw(B, g["state"], g["coef"], g["u0"], g["pred"], warm_in=g["wa"] if t > 0 else None,
                             status=g["stt"], iters=g["it"], warm_out=g["wb"], stream=sp)
                sv.warm_shift(B, g["wb"], g["wa"], stream=sp)
                sv.poststep_raw(B, g["u0"], g["vel"], None, g["cmd"], stream=sp)
                # plant (simulation): unicycle driven by the command
                sv.plant_step_raw(B, g["cmd"], g["pose"], g["vel"], stream=sp)
                if trace:
                    g["cte"].append(g["ce"][0].clone()); g["iters"].append(g["it"].clone()); g["conv"].append(g["stt"] == 1)
    for g in grp:
        torch.cuda.current_stream().wait_stream(g["st"])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    out = dict(loop_ms=ms, robot_ticks_per_s=R * T / (ms * 1e-3))
    if trace:
        out["cte"] = torch.cat([torch.stack(g["cte"]) for g in grp], dim=1).cpu().numpy()
        out["iters"] = torch.cat([torch.stack(g["iters"]) for g in grp], dim=1).cpu().numpy()
        out["conv"] = torch.cat([torch.stack(g["conv"]) for g in grp], dim=1).cpu().numpy()
    sv.close()
    return out


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 500
    cold = "--cold" in sys.argv
    K = 0
    if "--oracle-subset" in sys.argv:
        K = int(sys.argv[sys.argv.index("--oracle-subset") + 1])
    if "--device" in sys.argv:
        d = run_gpu_device(R, T, trace=False)
        dt_ = run_gpu_device(R, min(T, 100), trace=True)
        print(json.dumps(dict(robots=R, ticks=T, mode="device-resident loop, 8 streams", loop_ms=d["loop_ms"],
                              robot_ticks_per_s=d["robot_ticks_per_s"], mean_abs_cte_100=float(np.abs(dt_["cte"]).mean()),
                              median_abs_cte_100=float(np.median(np.abs(dt_["cte"]))),
                              mean_iters_100=float(dt_["iters"].mean()), converged_100=float(dt_["conv"].mean()))))
        return
    g = run_gpu(R, T, warm=not cold)
    res = dict(robots=R, ticks=T, warm=not cold, solve_s=g["solve_s"], solves_per_s=R * T / g["solve_s"],
               mean_abs_cte=float(np.abs(g["cte"]).mean()), max_abs_cte=float(np.abs(g["cte"]).max()),
               mean_abs_etheta=float(np.abs(g["eth"]).mean()), mean_iters=float(g["iters"].mean()),
               converged_fraction=float(g["conv"].mean()),
               mean_abs_cte_last100=float(np.abs(g["cte"][-100:]).mean()))
    res["per_track"] = per_track(g["kind"], g["cte"], g["dist"])
    if K > 0:
        # the checker (cold-started oracle loop = the reference loop) lives with the tests
        from tests.closed_loop_ref import run_oracle
        o = run_oracle(K, T)
        gk = run_gpu(K, T, warm=not cold)
        res["per_track_subset_gpu"] = per_track(gk["kind"], gk["cte"], gk["dist"])
        res["per_track_subset_oracle"] = per_track(o["kind"], o["cte"], o["dist"])
        res["oracle_subset"] = dict(robots=K, mean_abs_cte_oracle=float(np.abs(o["cte"]).mean()),
                                    mean_abs_cte_gpu=float(np.abs(gk["cte"]).mean()),
                                    max_trace_gap_cte=float(np.abs(np.abs(o["cte"]).mean(1) - np.abs(gk["cte"]).mean(1)).max()),
                                    mean_iters_oracle=float(o["iters"].mean()), mean_iters_gpu=float(gk["iters"].mean()))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
