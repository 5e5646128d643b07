"""tests/golden/make_golden.py -- generates the committed golden fixtures from the REFERENCE itself.

Run in the build container (needs /root/reference and oracle/_ref/libmpc_ref.so):
    python tests/golden/make_golden.py
Writes:
  fg_eval_golden.json   f, grad f, g, Jacobian and Lagrangian-Hessian nonzeros of the reference's
                        unmodified FG_eval (mpc_ros/src/mpc_planner.cpp:102-217) differentiated by its
                        vendored CppAD, at seeded points (SURVEY section 8c recipe);
  solve_golden.json     MPC::Solve results of the reference class (solver inside: oracle/ipm.c, the
                        stand-in for Ipopt 3.12.8 which is not installed) for seeded problems;
  hs071_golden.json     the HS071 known answer the reference ships
                        (assets/document/example/CppAD_Ipopt.cpp:146-150) and what the stand-in returns.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.oracle_py import Reference, YAML_DEFAULT, CFG_DEFAULT  # noqa: E402
from tests.problems import mild  # noqa: E402


def sparse(M, tol=0.0):
    idx = np.argwhere(np.abs(M) > tol)
    return [[int(i), int(j), float(M[i, j])] for i, j in idx]


def main():
    out = []
    cases = [
        # SURVEY 8c golden recipe: f = 4796.9161862598567, nnz 424 / 210
        dict(pm=dict(DT=.1, STEPS=20, REF_V=.5, REF_CTE=0, REF_ETHETA=0, W_CTE=100, W_EPSI=7, W_V=1000, W_ANGVEL=100,
                     W_A=50, W_DANGVEL=3, W_DA=10, ANGVEL=1.5, MAXTHR=1, BOUND=1e3),
             state=[0.03, 0, 0.02, 0.3, 0.05, -0.1], coeffs=[0.05, -0.1, 0.02, 0.003], lam="cos"),
        dict(pm=dict(YAML_DEFAULT), state=[0, 0, 0, 0.3, 0.05, -0.1], coeffs=[0.05, -0.1, 0.02, 0.003], lam="one"),
        dict(pm=dict(CFG_DEFAULT), state=[0.1, -0.2, 0.3, 0.7, -0.4, 0.5], coeffs=[-0.4, 0.6, -0.2, 0.05], lam="cos"),
        dict(pm=dict(YAML_DEFAULT, STEPS=7), state=[0, 0, 0, 0.1, 0.2, 0.3], coeffs=[0.2, 0.3, 0.1, -0.02], lam="cos"),
    ]
    for cs in cases:
        pm = cs["pm"]; N = int(pm["STEPS"]); n = 8 * N - 2; m = 6 * N
        R = Reference(pm)
        start = np.zeros(n)
        for k in range(6):
            start[k * N] = cs["state"][k]
        x = start + 0.01 * np.sin(1 + np.arange(n))
        lam = 0.5 + 0.1 * np.cos(0.7 * np.arange(m)) if cs["lam"] == "cos" else np.ones(m)
        r = R.fg_eval(cs["state"], cs["coeffs"], x, lam, 1.0)
        out.append(dict(params=pm, state=cs["state"], coeffs=cs["coeffs"], x=x.tolist(), lam=lam.tolist(),
                        f=r["f"], grad=r["grad"].tolist(), g=r["g"].tolist(), jac=sparse(r["J"]),
                        hess_lower=sparse(np.tril(r["H"])), nnz_jac=r["nnz_jac"], nnz_hess=r["nnz_hess"]))
    json.dump(out, open(os.path.join(HERE, "fg_eval_golden.json"), "w"))

    sol = []
    R = Reference(YAML_DEFAULT)
    R.set_cpu_time_override(100.0)
    state, coeffs = mild(1234, 12)
    for i in range(state.shape[1]):
        r = R.solve(state[:, i], coeffs[:, i])
        sol.append(dict(params=dict(YAML_DEFAULT), state=state[:, i].tolist(), coeffs=coeffs[:, i].tolist(),
                        u0=r["u0"].tolist(), pred=r["pred"].tolist(), obj=r["obj"], status=r["status"],
                        iters=r["iters"], kkt_error=r["kkt_error"]))
    R2 = Reference(dict(CFG_DEFAULT, W_DA=0.0))   # cfg weights without the rate term (GPU path scope)
    R2.set_cpu_time_override(100.0)
    state, coeffs = mild(99, 6)
    for i in range(state.shape[1]):
        r = R2.solve(state[:, i], coeffs[:, i])
        sol.append(dict(params=dict(CFG_DEFAULT, W_DA=0.0), state=state[:, i].tolist(), coeffs=coeffs[:, i].tolist(),
                        u0=r["u0"].tolist(), pred=r["pred"].tolist(), obj=r["obj"], status=r["status"],
                        iters=r["iters"], kkt_error=r["kkt_error"]))
    json.dump(sol, open(os.path.join(HERE, "solve_golden.json"), "w"))

    h = R.hs071(1e-8)
    json.dump(dict(known_x=[1.000000, 4.743000, 3.82115, 1.379408], known_zl0=1.087871, known_tol=1e-6,
                   standin=dict(status=h["status"], x=h["x"].tolist(), zl=h["zl"].tolist(), zu=h["zu"].tolist(),
                                obj=h["obj"], iters=h["iters"])),
              open(os.path.join(HERE, "hs071_golden.json"), "w"))
    print("wrote golden fixtures:", len(out), "fg_eval cases,", len(sol), "solves")


if __name__ == "__main__":
    main()
