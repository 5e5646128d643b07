#!/bin/bash
# round 2, pass 24: why does bench.py prefer 64 x 8 where gpu_sat.py prefers 128 x 4?  (number of data sets)
O=gpurun_out; mkdir -p $O
run() { echo "== $*"; timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 8 "$@" 2>&1 | tail -1 | python -c "
import json,sys
b=json.loads(sys.stdin.read()); print('value %.2f M/s  e2e %.2f M/s  frac %.3f  conc %.1f' % (b['value']/1e6, b['e2e']['value']/1e6, b['roofline']['frac'], b['roofline']['concurrency']))"; }
{
run --streams 128 --max-ctas 4 --sets 8
run --streams 128 --max-ctas 4 --sets 256
run --streams 64 --max-ctas 8 --sets 8
run --streams 64 --max-ctas 8 --sets 256
} > $O/r2y_sets.txt 2>&1
cat $O/r2y_sets.txt
