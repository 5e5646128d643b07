#!/bin/bash
# round 2, pass 8: why bench.py's value (rotating over > L2 of data) is below the L2-resident saturation run: harness sweep
O=gpurun_out; mkdir -p $O
run() { echo "== $*"; timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 8 "$@" 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read()); print('value %.2f M/s  e2e %.2f M/s  frac %.3f  launch_ms_event %.2f conc %.1f' % (b['value']/1e6, b['e2e']['value']/1e6, b['roofline']['frac'], b['roofline']['launch_ms_event_avg'], b['roofline']['concurrency']))"; }
{
run
run --sets 8
run --max-ctas 8
run --max-ctas 2
run --streams 256 --max-ctas 2
run --streams 64 --max-ctas 8
} > $O/r2h_sweep.txt 2>&1
cat $O/r2h_sweep.txt
