#!/bin/bash
# round 2, pass 22: dual-group kernel for the rate-penalty variant (28 lanes)
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 > $O/r2v_pytest.log; cat $O/r2v_pytest.log
python - <<'PY' > $O/r2v_rate.txt 2>&1
import sys, numpy as np
sys.path.insert(0, '.')
from mpc_ros_b200 import capi
from bench import gen_py
from oracle.oracle_py import CFG_DEFAULT
n = 16384
g = gen_py.problems(20261018 + 12, n)
prm = capi.params_from_map(CFG_DEFAULT, capi.yaml_default_params()); prm.delay_mode = 0; prm.max_iter = 100
for dual in (1, 0, 1, 0):
    sv = capi.Solver(prm, n, 0); sv.set_option("dual_groups", dual); sv.set_option("max_ctas", 64)
    coeffs, state = sv.prestep(g["wx"], g["wy"], g["pose"], g["vel"])
    best = 1e9
    for _ in range(3):
        r = sv.solve(state, coeffs); best = min(best, sv.last_kernel_seconds)
    print("cfg weights, 16,384 problems on 64 CTAs, dual %d: kernel %.3f ms, converged %d, mean iters %.3f" % (dual, best * 1e3, int((r["status"] == 1).sum()), r["iters"].mean()))
    sv.close()
PY
cat $O/r2v_rate.txt
