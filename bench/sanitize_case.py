"""bench/sanitize_case.py -- small cold + warm + rate-penalty solves for compute-sanitizer runs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mpc_ros_b200 import capi
from bench import gen_py

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    prm = capi.yaml_default_params(); prm.delay_mode = 0
    sv = capi.Solver(prm, B, 0)
    sv.set_option("max_ctas", 2)          # lanes refill from the queue
    sv.set_option("problems_per_cta", 32)
    g = gen_py.problems(20261020, B)
    coeffs, state = sv.prestep(g["wx"], g["wy"], g["pose"], g["vel"])
    out = sv.solve(state, coeffs)
    print("cold: converged", int((out["status"] == 1).sum()), "of", B)
    dev = torch.device("cuda:0"); N = 20
    ws = capi.lib().mpc_b200_warm_size(N)
    f64 = dict(dtype=torch.float64, device=dev)
    ds = torch.from_numpy(state).to(dev); dc = torch.from_numpy(coeffs).to(dev)
    u = torch.zeros((2, B), **f64); pr = torch.zeros((3 * N, B), **f64); wo = torch.zeros((ws, B), **f64); w2 = torch.zeros((ws, B), **f64)
    st = torch.zeros(B, dtype=torch.int32, device=dev)
    sv.solve_raw(B, ds, dc, u, pr, status=st, warm_out=wo); sv.warm_shift(B, wo, w2)
    sv.solve_raw(B, ds, dc, u, pr, warm_in=w2, status=st); torch.cuda.synchronize()
    print("warm: converged", int((st == 1).sum().item()))
    sv.close()
    prm.w_accel_d = 10.0; prm.w_angvel_d = 3.0
    sv = capi.Solver(prm, B, 0); sv.set_option("max_ctas", 2)
    out = sv.solve(state, coeffs)
    print("rate: converged", int((out["status"] == 1).sum()))
    sv.close()
main()
