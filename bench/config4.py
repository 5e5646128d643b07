"""bench/config4.py -- BASELINE config 4: long horizon N = 100, batch 16,384 (shared-memory staging stress)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mpc_ros_b200 import capi
from bench import gen_py

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    prm = capi.yaml_default_params(); prm.mpc_steps = N; prm.delay_mode = 0
    sv = capi.Solver(prm, B, 0)
    g = gen_py.problems(20261018 + 4, B)
    coeffs, state = sv.prestep(g["wx"], g["wy"], g["pose"], g["vel"])
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter(); out = sv.solve(state, coeffs); t1 = time.perf_counter()
        best = min(best, sv.last_kernel_seconds)
    conv = (out["status"] == 1) & (out["kkt"] <= 1e-8)
    it = out["iters"]
    flop_iter = (N - 1) * (842.67 + 392 + 80) + N * (100 + 45)
    print(json.dumps(dict(config=4, batch=B, mpc_steps=N, kernel_ms=best * 1e3, converged=int(conv.sum()),
                          converged_fraction=float(conv.mean()), solves_per_s=float(conv.sum() / best),
                          mean_iters=float(it[conv].mean()), max_iters=int(it.max()),
                          algorithmic_tflops=float(flop_iter * it.sum() / best / 1e12),
                          status_hist={int(k): int(v) for k, v in zip(*np.unique(out["status"], return_counts=True))})))
    sv.close()
main()
