"""bench/config4_prof.py -- per-phase cycle profile of the solve kernel on BASELINE config 4 (N = 100); needs the NMPC_PROFILE
build: MPC_B200_LIB=mpc_ros_b200/lib/libmpc_b200_prof.so python bench/config4_prof.py [batch] [N]."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mpc_ros_b200 import capi
from bench import gen_py

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    L = capi.lib()
    prm = capi.yaml_default_params(); prm.mpc_steps = N; prm.delay_mode = 0
    sv = capi.Solver(prm, B, 0)
    g = gen_py.problems(20261018 + 4, B)
    coeffs, state = sv.prestep(g["wx"], g["wy"], g["pose"], g["vel"])
    out = sv.solve(state, coeffs)
    has = hasattr(L, "mpc_b200_debug_profile") and L.mpc_b200_debug_profile(sv._h, None)
    out = sv.solve(state, coeffs)
    ks = sv.last_kernel_seconds
    conv = (out["status"] == 1) & (out["kkt"] <= 1e-8)
    res = dict(batch=B, mpc_steps=N, kernel_ms=ks * 1e3, solves_per_s=float(conv.sum() / ks), mean_iters=float(out["iters"][conv].mean()),
               iters_hist=np.bincount(np.minimum(out["iters"], 100) // 10).tolist())
    if has:
        buf = (C.c_longlong * 1024)()
        L.mpc_b200_debug_profile(sv._h, buf)
        names = ["-", "refill", "P3_apply_coeffs", "P4_backward", "P4_forward", "P5_step", "P6", "P1_eval", "P2_rest", "P6_adjoint", "P2_ctrl_decide"]
        res["avg_cycle"] = buf[1001] / max(1, buf[1000]); res["cycles_total"] = int(buf[1000])
        res["busy_lanes"] = buf[1002] / max(1, buf[1000])
        res["per_cycle"] = {names[i]: round(buf[1008 + i] / max(1, buf[1000])) for i in range(1, 11)}
    print(json.dumps(res))
    sv.close()
main()
