#!/bin/bash
# profiles/evidence_run.sh TAG -- everything profiles/make_summary.py TAG needs, in one gpurun call:
#   gpurun --timeout 2400 -- 'bash profiles/evidence_run.sh r1'
# (ncu passes run only after the same command has exited 0 without ncu; numbers printed under ncu are not used.)
T=${1:-r1}; O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py > $O/bench_default_$T.json 2> $O/bench_default_$T.err
CMD="python bench.py --steps 40 --warmup 8 --streams 8 --no-cpu-baseline --e2e-steps 16 --e2e-threads 2 --e2e-inflight 4"
$CMD > $O/plain_$T.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $O/launches_$T.csv $CMD > $O/ncu_list.log 2>&1
$CMD > $O/plain_${T}b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:nmpc_solve -s 10 -c 1 \
    -f -o $O/prof_${T}_solve $CMD > $O/ncu_full.log 2>&1
tail -1 $O/ncu_full.log
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_$T.json
mpc_ros_b200/lib/mpc_bench latency 10000 > $O/latency_$T.json
python bench/config4.py 16384 100 > $O/config4_$T.json
python bench/closed_loop.py 1024 500 --oracle-subset 32 > $O/config5_$T.json
python bench/closed_loop.py 1024 500 --device > $O/config5_device_$T.json
# (config 3, multi-GPU: gpurun --gpus 4 -- torchrun ... bench/config3.py; see profiles/r1_config3.json)
python tests/parity_sweep.py 16384 > $O/parity_sweep_$T.json 2> $O/parity_sweep.err
cut -c1-300 $O/bench_default_$T.json
