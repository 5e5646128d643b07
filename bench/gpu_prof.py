"""bench/gpu_prof.py -- per-phase cycle profile of the solve kernel (needs the NMPC_PROFILE build:
MPC_B200_LIB=mpc_ros_b200/lib/libmpc_b200_prof.so python bench/gpu_prof.py [batch])."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from mpc_ros_b200 import capi  # noqa: E402
from bench import gen_py  # noqa: E402

NAMES = ["-", "refill", "P3_apply_coeffs", "P4_backward", "P4_forward", "P5_step", "P6_adjoint_stepsize",
         "P1_eval", "P2_decide", "-"]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    same = len(sys.argv) > 2 and sys.argv[2] == "same"
    L = capi.lib()
    L.mpc_b200_measure_fp64_peak(0, 200000)    # ~0.3 s of DFMA: clocks up
    prm = capi.yaml_default_params()
    g = gen_py.problems(20261018 + 2, B)
    s = capi.Solver(prm, B, 0)
    coeffs, cte, eth = s.polyfit(g["wx"], g["wy"], g["pose"])
    state = np.zeros((6, B)); state[3] = g["vel"][0]; state[4] = cte; state[5] = eth
    if same:      # B copies of problem 0 in one CTA: every lane does the same thing, clean per-phase costs
        state = np.repeat(state[:, :1], B, axis=1).copy(); coeffs = np.repeat(coeffs[:, :1], B, axis=1).copy()
        s.set_option("problems_per_cta", 32)
    has_prof = hasattr(L, "mpc_b200_debug_profile") and L.mpc_b200_debug_profile(s._h, None)
    for rep in range(5):
        out = s.solve(state, coeffs)
        ks = s.last_kernel_seconds
    res = dict(batch=B, kernel_us=ks * 1e6, iters_mean=float(out["iters"].mean()), iters_max=int(out["iters"].max()),
               iters_first32_max=int(out["iters"][:32].max()))
    if has_prof:
        buf = (C.c_longlong * 1024)()
        L.mpc_b200_debug_profile(s._h, buf)
        cyc = {NAMES[i]: int(buf[i]) for i in range(10)}
        res["cta0_cycles"] = cyc
        res["cta0_total_cycles"] = sum(cyc.values())
        t0 = buf[16]
        for cy in (1, 2, 3):
            res["trace_ctrl_%d" % cy] = [int(buf[16 + 16 * cy + i] - t0) for i in range(11)]
            res["trace_stage_%d" % cy] = [int(buf[512 + 16 * cy + i] - t0) for i in range(10)]
    print(json.dumps(res, indent=1))
    s.close()


if __name__ == "__main__":
    main()
