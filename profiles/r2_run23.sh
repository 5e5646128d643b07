#!/bin/bash
# round 2, pass 23: does the nvidia-smi clock sampler perturb the timed region?
O=gpurun_out; mkdir -p $O
run() { echo "== $*"; timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 8 "$@" 2>&1 | tail -1 | python -c "
import json,sys
b=json.loads(sys.stdin.read()); print('value %.2f M/s  e2e %.2f M/s  frac %.3f  conc %.1f  clock samples %s' % (b['value']/1e6, b['e2e']['value']/1e6, b['roofline']['frac'], b['roofline']['concurrency'], b['clocks'].get('samples')))"; }
{
run --streams 64 --max-ctas 8 --clock-period-ms 50
run --streams 64 --max-ctas 8 --clock-period-ms 500
run --streams 128 --max-ctas 4 --clock-period-ms 50
run --streams 128 --max-ctas 4 --clock-period-ms 500
} > $O/r2x_clock.txt 2>&1
cat $O/r2x_clock.txt
