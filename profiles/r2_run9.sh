#!/bin/bash
# round 2, pass 9: harness sweep around 64 streams x 8 CTAs
O=gpurun_out; mkdir -p $O
run() { echo "== $*"; timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 8 "$@" 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read()); print('value %.2f M/s  e2e %.2f M/s  frac %.3f  launch_ms_event %.2f conc %.1f' % (b['value']/1e6, b['e2e']['value']/1e6, b['roofline']['frac'], b['roofline']['launch_ms_event_avg'], b['roofline']['concurrency']))"; }
{
run --streams 64 --max-ctas 8
run --streams 32 --max-ctas 8
run --streams 48 --max-ctas 8
run --streams 64 --max-ctas 6
run --streams 32 --max-ctas 16
run --streams 64 --max-ctas 12
run --streams 96 --max-ctas 8
run --streams 64 --max-ctas 8 --e2e-threads 4 --e2e-inflight 128
run --streams 64 --max-ctas 8 --e2e-threads 2 --e2e-inflight 64 --e2e-max-ctas 8
} > $O/r2i_sweep.txt 2>&1
cat $O/r2i_sweep.txt
