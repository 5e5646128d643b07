#!/bin/bash
# round 2, pass 21: hardware queues (CUDA_DEVICE_MAX_CONNECTIONS) x streams x CTAs per launch in bench.py
O=gpurun_out; mkdir -p $O
run() { echo "== conn $1: ${@:2}"; CUDA_DEVICE_MAX_CONNECTIONS=$1 timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 8 "${@:2}" 2>&1 | tail -1 | python -c "
import json,sys
b=json.loads(sys.stdin.read()); print('value %.2f M/s  e2e %.2f M/s  frac %.3f  launch_ms_event %.2f conc %.1f' % (b['value']/1e6, b['e2e']['value']/1e6, b['roofline']['frac'], b['roofline']['launch_ms_event_avg'], b['roofline']['concurrency']))"; }
{
run 32 --streams 64 --max-ctas 8
run 8 --streams 64 --max-ctas 8
run 8 --streams 128 --max-ctas 4
run 16 --streams 128 --max-ctas 4
run 4 --streams 128 --max-ctas 4
run 8 --streams 256 --max-ctas 2
} > $O/r2u_conn.txt 2>&1
cat $O/r2u_conn.txt
