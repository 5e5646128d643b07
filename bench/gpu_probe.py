"""bench/gpu_probe.py -- first-contact GPU measurements: FP64 peak / issue probes, parity spot check,
kernel time vs batch.  Writes gpurun_out/probe.json."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from mpc_ros_b200 import capi  # noqa: E402
from bench import gen_py  # noqa: E402


def main():
    L = capi.lib()
    res = {}
    res["fp64_peak_tflops"] = L.mpc_b200_measure_fp64_peak(0, 20000)
    probes = {}
    for ilp in (1, 2, 4, 8):
        for lanes in (32, 16, 8, 1):
            probes["ilp%d_lanes%d" % (ilp, lanes)] = L.mpc_b200_debug_fp64_probe(0, ilp, lanes, 20000)
    res["dfma_cycles_per_instr"] = probes
    print(json.dumps(res, indent=1))
    prm = capi.yaml_default_params()
    timings = {}
    for B in (1, 32, 148, 1024, 4096, 4736, 8192, 16384, 65536):
        g = gen_py.problems(20261018 + 2, B)
        s = capi.Solver(prm, B, 0)
        coeffs, cte, eth = s.polyfit(g["wx"], g["wy"], g["pose"])
        state = np.zeros((6, B)); state[3] = g["vel"][0]; state[4] = cte; state[5] = eth
        ks = []
        for rep in range(4):
            t0 = time.perf_counter()
            out = s.solve(state, coeffs)
            t1 = time.perf_counter()
            ks.append((s.last_kernel_seconds, t1 - t0))
        it = out["iters"]
        timings[B] = dict(kernel_s=min(k for k, _ in ks), e2e_s=min(e for _, e in ks),
                          converged=int((out["status"] == 1).sum()), iters_mean=float(it.mean()), iters_max=int(it.max()),
                          status_hist={int(k): int(v) for k, v in zip(*np.unique(out["status"], return_counts=True))})
        print(B, timings[B], flush=True)
        s.close()
    res["timings"] = timings
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
