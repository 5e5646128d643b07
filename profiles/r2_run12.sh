#!/bin/bash
# round 2, pass 12: backtracking fast path of the dual-group kernel -- agreement, one-shot config 3, saturated throughput
O=gpurun_out; mkdir -p $O
{
timeout 120 python bench/dual_check.py 20000 8 2>&1 | tail -9
timeout 300 python -m pytest tests -m gpu -q -x -k "dual or special or sweep_config3 or warm" 2>&1 | tail -2
echo "== config 3 one-shot, 1 GPU"; timeout 300 mpc_ros_b200/lib/mpc_bench multi 1 65536 7 | tail -1 | cut -c1-330
echo "== sat dual"; timeout 200 python bench/gpu_sat.py 4096 128 3000 4 0 1 1 2>&1 | tail -2
} > $O/r2l_fast.txt 2>&1
cat $O/r2l_fast.txt
