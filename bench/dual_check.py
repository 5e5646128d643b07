"""bench/dual_check.py -- the dual-group kernel (nmpc_kernel_dual.cuh) against the single-group kernel on the same batch:
status, iteration counts, first controls, objective and predicted states must agree bit for bit (the two kernels run the
same phase functions and add the partial sums in the same order).  argv: [problems [max_ctas [warm]]]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mpc_ros_b200 import capi
from bench import gen_py


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    maxc = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    g = gen_py.problems(20261018 + 3, n)
    prm = capi.yaml_default_params(); prm.delay_mode = 0; prm.max_iter = 100
    out = {}
    for dual in (1, 0):
        sv = capi.Solver(prm, n, 0)
        sv.set_option("dual_groups", dual)
        sv.set_option("problems_per_cta", 32)
        if maxc: sv.set_option("max_ctas", maxc)
        coeffs, state = sv.prestep(g["wx"], g["wy"], g["pose"], g["vel"])
        r = sv.solve(state, coeffs)
        r2 = sv.solve(state, coeffs)
        ms = sv.last_kernel_seconds * 1e3
        out[dual] = (r, ms)
        print("dual %d: kernel %.3f ms, converged %d of %d, mean iters %.3f, repeat identical %s" %
              (dual, ms, int((r["status"] == 1).sum()), n, r["iters"].mean(), bool((r["u0"] == r2["u0"]).all())))
        sv.close()
    a, b = out[1][0], out[0][0]
    for key in ("status", "iters", "u0", "obj", "pred", "kkt"):
        same = np.array_equal(a[key], b[key])
        print("%-6s identical: %s%s" % (key, same, "" if same else "  max abs diff %.3e" % np.abs(a[key].astype(float) - b[key].astype(float)).max()))
    assert np.array_equal(a["status"], b["status"]) and np.array_equal(a["u0"], b["u0"]) and np.array_equal(a["pred"], b["pred"])
    print("OK")


main()
