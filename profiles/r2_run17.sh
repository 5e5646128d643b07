#!/bin/bash
# round 2, pass 17: harness sweep with the dual-group kernel
O=gpurun_out; mkdir -p $O
run() { echo "== $*"; timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 8 "$@" 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read()); print('value %.2f M/s  e2e %.2f M/s  frac %.3f  launch_ms_event %.2f conc %.1f' % (b['value']/1e6, b['e2e']['value']/1e6, b['roofline']['frac'], b['roofline']['launch_ms_event_avg'], b['roofline']['concurrency']))"; }
{
run --streams 64 --max-ctas 8
run --streams 128 --max-ctas 4
run --streams 96 --max-ctas 6
run --streams 64 --max-ctas 8 --depth 3
run --streams 48 --max-ctas 12
run --streams 64 --max-ctas 6
} > $O/r2q_sweep.txt 2>&1
cat $O/r2q_sweep.txt
