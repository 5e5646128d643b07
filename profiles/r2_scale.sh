#!/bin/bash
# profiles/r2_scale.sh N -- bench.py on N GPUs of one box the way the driver launches it, BASELINE config 3 (strong) through the
# C++ harness, and the copy-only probe (what the host's memory / PCIe fabric can feed without any solve)
N=${1:-8}; O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c; nproc
if [ "$N" = 1 ]; then python bench.py --gpus 1 > $O/r2_bench_n$N.json 2> $O/r2_bench_n$N.err
else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > $O/r2_bench_n$N.json 2> $O/r2_bench_n$N.err; fi
tail -2 $O/r2_bench_n$N.err; cut -c1-250 $O/r2_bench_n$N.json
for g in 1 2 4 8; do if [ $g -le $N ]; then timeout 300 mpc_ros_b200/lib/mpc_bench multi $g 65536 7 | tail -1; fi; done > $O/r2_config3_upto$N.jsonl
cat $O/r2_config3_upto$N.jsonl | cut -c1-330
for g in 1 $N; do
if [ "$g" = 1 ]; then python bench/probe/copy_scale.py 4000; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29512 bench/probe/copy_scale.py 4000 2>/dev/null; fi
done > $O/r2_copy_scale_n$N.jsonl
cat $O/r2_copy_scale_n$N.jsonl
