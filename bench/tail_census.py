"""bench/tail_census.py -- CPU study of the tail of a one-shot batch (BASELINE config 3): which problems take how many
global cycles of the solve kernel, and what the cycles are (Newton iterations, least-squares start, backtracking trials,
second-order corrections, resumed line searches, inertia-correction retries).  Runs the kernel's phase functions on the
host emulator (tests/emu, test infrastructure) over the config-3 problem set and the oracle (the Ipopt restatement) on the
slowest problems.  No GPU needed:    python bench/tail_census.py [problems] > profiles/rN_tail_census.txt"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from bench import gen_py  # noqa: E402
from oracle.oracle_py import Oracle, YAML_DEFAULT  # noqa: E402
from tests import test_emu  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    g = gen_py.problems(20261018 + 3, n)
    orc = Oracle()
    state = np.zeros((6, n)); coeffs = np.zeros((4, n))
    for i in range(n):
        c, cte, eth = orc.prestep(g["wx"][:, i], g["wy"][:, i], *g["pose"][:, i])
        coeffs[:, i] = c; state[3, i] = g["vel"][0, i]; state[4, i] = cte; state[5, i] = eth
    r = test_emu.emu_solve(YAML_DEFAULT, state, coeffs, PB=32, max_iter=100)
    ip = C.POINTER(C.c_int)
    ages = np.zeros(n, dtype=np.int32); test_emu._EMU.nmpc_emu_last_ages(ages.ctypes.data_as(ip), n)
    k = np.zeros(4 * n, dtype=np.int32); test_emu._EMU.nmpc_emu_last_kinds(k.ctypes.data_as(ip), n); k = k.reshape(n, 4)
    print("config-3 problem set, %d problems, YAML weights, max_iter 100 (kernel logic on the host emulator)" % n)
    print("status histogram:", {int(a): int(b) for a, b in zip(*np.unique(r["status"], return_counts=True))})
    print("global cycles per problem: mean %.2f, quantiles 50/90/99/99.9/99.99/100 %%: %s" %
          (ages.mean(), np.quantile(ages, [.5, .9, .99, .999, .9999, 1]).tolist()))
    print("all cycles %d = iterations %d + least-squares starts %d + backtracking trials %d + second-order corrections %d"
          " + resumed line searches %d + inertia retries %d   (restoration steps are counted with the corrections)" %
          (ages.sum(), r["iters"].sum(), n, k[:, 0].sum(), k[:, 1].sum(), k[:, 2].sum(), k[:, 3].sum()))
    for thr in (24, 50, 100):
        m = ages > thr
        print("problems with more than %3d cycles: %4d (%.3f %%), %.1f %% of all cycles; their cycles: iterations %d, backtracks %d,"
              " corrections %d, resumes %d, inertia retries %d" %
              (thr, m.sum(), 100.0 * m.mean(), 100.0 * ages[m].sum() / ages.sum(), r["iters"][m].sum(), k[m, 0].sum(),
               k[m, 1].sum(), k[m, 2].sum(), k[m, 3].sum()))
    print("track kinds of the >100-cycle problems (0 infinity, 1 epitrochoid, 2 square):",
          {int(a): int(b) for a, b in zip(*np.unique(g["kind"][ages > 100], return_counts=True))})
    opt = orc.default_options(); opt.max_iter = 100
    print("\nthe 25 slowest problems: kernel logic vs the oracle (Ipopt's algorithm on the CPU)")
    print("%7s %7s %6s %6s | %10s %10s %8s   %s" % ("problem", "cycles", "iters", "status", "oracle it", "oracle st", "restored", "path polynomial"))
    for i in np.argsort(-ages)[:25]:
        o = orc.solve(YAML_DEFAULT, state[:, i], coeffs[:, i], opt)
        print("%7d %7d %6d %6d | %10d %10d %8d   %s" % (i, ages[i], r["iters"][i], r["status"][i], o["iters"], o["status"],
                                                        o.get("n_resto", 0), np.array2string(coeffs[:, i], precision=1)))


main()
