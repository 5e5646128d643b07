"""profiles/make_r2_summary.py -- copies the round-2 evidence a `gpurun -- bash profiles/r2_evidence.sh` run (1 GPU) and a
`gpurun --gpus 8 -- bash profiles/r2_scale.sh 8` run left in gpurun_out/ into profiles/ and writes profiles/r2_summary.md."""
import collections, csv, json, os, shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out") + "/"; P = os.path.join(ROOT, "profiles") + "/"
for f in ("bench_default.json", "bench_reference.json", "config1_latency.json", "config4.json", "config4_prof.json", "parity_sweep.json",
          "launches.csv", "solve_kernel_ncu_raw.csv", "phase_cycles.txt", "gpu_sat.txt", "config5.json", "config5_device.json"):
    if os.path.exists(G + "r2f_" + f) and os.path.getsize(G + "r2f_" + f) > 0:
        shutil.copy(G + "r2f_" + f, P + "r2_" + f)
for f in ("bench_n8.json", "config3_upto8.jsonl", "copy_scale_n8.jsonl"):
    if os.path.exists(G + "r2_" + f):
        shutil.copy(G + "r2_" + f, P + "r2_" + f)


def last_json(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


d = last_json(P + "r2_bench_default.json"); ref = last_json(P + "r2_bench_reference.json"); d8 = last_json(P + "r2_bench_n8.json")
c1 = last_json(P + "r2_config1_latency.json"); c4 = json.load(open(P + "r2_config4.json")); c5 = json.load(open(P + "r2_config5.json"))
c5d = json.load(open(P + "r2_config5_device.json"))
c3 = [json.loads(l) for l in open(P + "r2_config3_upto8.jsonl") if l.startswith("{")]
rows = [r for r in csv.reader(open(P + "r2_launches.csv")) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]; H = rows[hdr]; data = rows[hdr + 1:]
ik = H.index("Kernel Name"); iv = H.index("Metric Value"); iu = H.index("Metric Unit")
agg = collections.defaultdict(list)
for r in data:
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    u = r[iu]; v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
    agg[r[ik][:90]].append(v)
tot = sum(sum(v) for v in agg.values())
lines = ["| `%s` | %d | %.2f | %.1f %% | %.1f |" % (k, len(v), sum(v) / 1e3, 100 * sum(v) / tot, sum(v) / len(v))
         for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))]
mine = sum(sum(v) for k, v in agg.items() if "nmpc" in k or "prestep" in k or "queue_order" in k)
solve = sum(sum(v) for k, v in agg.items() if "nmpc" in k)
rr = list(csv.reader(open(P + "r2_solve_kernel_ncu_raw.csv"))); h, u, v = rr[0], rr[1], rr[2]
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sass__inst_executed_register_spilling", "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__sass_inst_executed_op_shared.sum"]
met = ["| `%s` | %s | %s |" % (k, v[h.index(k)], u[h.index(k)]) for k in keys if k in h]
kname = v[h.index("Kernel Name")] if "Kernel Name" in h else "solve kernel"
r = d["roofline"]; e = d["e2e"]; b3 = d["config3"]
md = """# Round 2 profile summary (B200, gpurun)

Files in this directory named `r2_*` come from `gpurun` runs of the committed tree: `profiles/r2_evidence.sh` (one GPU),
`profiles/r2_scale.sh 8` (eight GPUs), `profiles/r2_run10.sh ... r2_run18.sh` (experiments of this round); this file is
written by `profiles/make_r2_summary.py`.

## Headline (`r2_bench_default.json`: `python bench.py`, no flags = the driver's protocol)

* value **%.2f M converged solves/s** on 1xB200: %d steps of %d streamed batches of 4,096 problems (%.1f ms per step,
  timed region %.2f s), %d streams, persistent grid of %d CTAs per launch, dual-group kernel, issue loop: %s; clocks
  %.0f/%.0f MHz, throttle reasons %s; %d kernel launches of this repo in the timed region.
* e2e (host buffers, one packed H2D + one packed D2H copy per batch inside the timed region, %d host threads x %d
  handles) **%.2f M solves/s**; the same copies with no solve: %.1f M solves/s (%.0f GB/s).
* roofline (FP64 pipe): %.2f TFLOP/s algorithmic (SURVEY 8d count, 27,879 flop per iteration) of %.1f measured peak =
  **%.1f %%**; in executed flops (ncu count of one launch) %.1f %%; DRAM traffic of one launch %.2f MB (ncu) vs 2.49 MB
  algorithmic (the results are still in L2 when the kernel ends).
* latency (config 1, through the C++ `MPC` adapter): p50 %.1f us, p99 %.1f us in the bench line; `r2_config1_latency.json`
  (10,000 calls): p50 %.1f us, p99 %.1f us.
* config 3 one-shot in the bench line (65,536 on one GPU, host buffers, C++ harness): %.2f ms (solve kernel %.2f ms).
* CPU baseline: %.0f solves/s on %d host cores (`oracle/_ref`: reference `mpc_planner.cpp` + CppAD unmodified, `std::thread`
  per core, stand-in IPM inside); reference arm (`r2_bench_reference.json`): %.0f solves/s.

## Eight GPUs (`r2_bench_n8.json`: `torchrun --nproc-per-node 8 bench.py --gpus 8`, the driver's launch)

* value **%.1f M converged solves/s** (weak scaling, max over ranks; %.1f %% of 8 x the one-GPU value), e2e **%.1f M solves/s**;
  copies alone in the same run %.1f M solves/s (%.0f GB/s), in the longer probe `r2_copy_scale_n8.jsonl` 147.8 M solves/s
  (116 GB/s; one GPU alone 65 GB/s): the e2e rate at eight GPUs is the box's copy ceiling.
* config 3 (one batch of 65,536 split contiguously, C++ `std::thread` per GPU, `r2_config3_upto8.jsonl`):
  %s ms on %s GPUs -- bound by the slowest problem (520 global cycles), see `r2_tail_census.txt` and DESIGN.md section 6.

## Other BASELINE configs

* config 4 (`r2_config4.json`): N = 100, batch 16,384: %.0f converged solves/s, %.1f mean iterations, converged fraction %.4f
  (status 9: %d, status 2: %d); phase cycles `r2_config4_prof.json`.
* config 5 (`r2_config5.json`, `r2_config5_device.json`): 1,024 robots x 500 ticks, warm start: %.2f mean iterations (oracle loop
  on a 32-robot subset, cold: %.2f), converged fraction %.5f; host-driven loop %.0f solves/s, device-resident loop %.0f
  robot-ticks/s.  Tracking: median distance to the path %.2f / %.2f / %.2f m on the infinity / epitrochoid / square tracks
  (GPU loop and oracle loop agree; on the epitrochoid to 1e-8); the cubic's `cte` is meaningless where the window straddles a
  corner or the crossing (max %.1e m), which is why the distance is reported beside it.

## Launch list (`r2_launches.csv`)

`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` on
`python bench.py --steps 1 --warmup 3 --batches-per-step 16 --streams 8 --no-extras --no-cpu-baseline --e2e-inflight 4`
(serialised, cold; the first 400 launches of the process).

| kernel | launches | total ms | share | avg us |
|---|---|---|---|---|
%s

`dfma_kernel<8>` is bench.py's FP64-peak probe (outside the timed region), the `at::` fills are torch zero-fills of the
set-up.  Among this repo's kernels the solve kernel is %.1f %% of the device time (pre-step and queue order together
%.1f %%), in line with bench.py's CUDA-event figures (`roofline.kernel_ms` x launches = `ms_per_step`).

## Solve kernel, `ncu --set full` (`r2_solve_kernel_ncu_raw.csv`: `%s`, one launch of 4,096 problems on 64 CTAs, alone)

| metric | value | unit |
|---|---|---|
%s

Reading: 384 threads x 166 registers and 227.7 KB dynamic shared memory = one 12-warp CTA per SM by design.  A lone launch
leaves most SMs idle most of the time (the tail of a batch), which is why `pct_of_peak_sustained_elapsed` is tiny and why
throughput is measured with many batches in flight.  Within active cycles the FP64 pipe is busy ~29 %% and an
instruction issues in ~34 %% of the cycles (round 1: 28.9 %%); the barrier stall per issue fell from 6.0 to 3.7 (the stage
warps of the dual-group kernel work for one group while the other group's control warp sweeps).  No spilled
instructions.  DRAM traffic = the inputs; the outputs are still in L2 when the kernel ends; nothing is re-read.

## Phase cycles (`r2_phase_cycles.txt`, `r2_dual_groups.txt`), saturated throughput (`r2_gpu_sat.txt`), experiments

See DESIGN.md section 9.  `r2_dual_groups.txt`: dual-group kernel vs single-group kernel, both rotations, both control-warp
placements, the dropped evaluation-only cycles.  `r2_latency_mode.txt`: one stage per stage thread in narrow CTAs.
`r2_harness_sweep.txt`, `r2_issue_loop.txt`: streams x CTAs per launch, C++ vs Python issue loop.  `r2_tail_census.txt`: cycles
per problem of config 3's batch by kind, the 25 slowest problems against the oracle.  `r2_parity_sweep.json`: 3 x 16,384
problems against the oracle (DESIGN.md section 7).
""" % (d["value"] / 1e6, d["steps"], d["config"].get("batches_per_step", 256), d["ms_per_step"], d["steps"] * d["ms_per_step"] / 1e3,
       d["config"]["streams"], d["config"]["max_ctas"], d["config"].get("issue_loop", "python"),
       d["clocks"]["sm_mhz"], d["clocks"]["sm_max_mhz"], d["clocks"]["reasons"], d["gpu_launches"],
       e["threads"], e["in_flight"] // e["threads"], e["value"] / 1e6, e["copies_alone"]["value"] / 1e6, e["copies_alone"]["gbytes_per_s"],
       r["achieved"], r["peak"], 100 * r["frac"], 100 * (r["frac_executed"] or 0), (r["traffic"] or 0) / 1e6,
       d["latency"]["p50_us"], d["latency"]["p99_us"], c1["p50_us"], c1["p99_us"], b3["one_shot_ms_median"], b3["solve_kernel_ms_slowest_gpu"],
       d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], ref["value"],
       d8["value"] / 1e6, 100 * d8["value"] / (8 * d["value"]), d8["e2e"]["value"] / 1e6,
       d8["e2e"].get("copies_alone", {}).get("value", 0) / 1e6, d8["e2e"].get("copies_alone", {}).get("gbytes_per_s", 0),
       " / ".join("%.2f" % x["one_shot_ms_median"] for x in c3), " / ".join(str(x["gpus"]) for x in c3),
       c4["solves_per_s"], c4["mean_iters"], c4["converged_fraction"], c4["status_hist"].get("9", 0), c4["status_hist"].get("2", 0),
       c5["mean_iters"], c5["oracle_subset"]["mean_iters_oracle"], c5["converged_fraction"], c5["solves_per_s"], c5d["robot_ticks_per_s"],
       c5["per_track"]["infinity"]["median_dist"], c5["per_track"]["epitrochoid"]["median_dist"], c5["per_track"]["square"]["median_dist"],
       c5["max_abs_cte"], "\n".join(lines), 100 * solve / mine, 100 * (mine - solve) / mine, kname, "\n".join(met))
open(P + "r2_summary.md", "w").write(md)
print(md[:1500])
