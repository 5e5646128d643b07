#!/bin/bash
# round 2, pass 18: bench.py's device-resident leg issued from C++ (bench/issue_loop.cpp) vs the Python loop
O=gpurun_out; mkdir -p $O
run() { echo "== $*"; timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 8 "$@" 2>&1 | tail -1 | python -c "
import json,sys
b=json.loads(sys.stdin.read()); print('value %.2f M/s  e2e %.2f M/s  frac %.3f  launch_ms_event %.2f conc %.1f  %s' % (b['value']/1e6, b['e2e']['value']/1e6, b['roofline']['frac'], b['roofline']['launch_ms_event_avg'], b['roofline']['concurrency'], b['config'].get('issue_loop')))"; }
{
run --streams 64 --max-ctas 8
run --streams 64 --max-ctas 8 --python-issue
run --streams 128 --max-ctas 4
run --streams 128 --max-ctas 4 --python-issue
run --streams 96 --max-ctas 6
run --streams 128 --max-ctas 4 --depth 3
} > $O/r2r_issue.txt 2>&1
cat $O/r2r_issue.txt
