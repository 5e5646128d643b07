#!/bin/bash
# profiles/r2_scale_n.sh N -- only the bench line on N GPUs (the driver's launch), for the scaling series in profiles/
N=${1:-2}; O=gpurun_out; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --no-cpu-baseline > $O/r2_bench_n$N.json 2> $O/r2_bench_n$N.err
tail -1 $O/r2_bench_n$N.err; cut -c1-200 $O/r2_bench_n$N.json
