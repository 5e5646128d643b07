#!/bin/bash
# round 2, pass 10: dual-group kernel -- agreement with the single-group kernel, then saturated throughput A/B and phases
O=gpurun_out; mkdir -p $O
{
timeout 120 python bench/dual_check.py 8192 2>&1 | tail -12
timeout 120 python bench/dual_check.py 20000 8 2>&1 | tail -10
echo "== sat dual"; timeout 200 python bench/gpu_sat.py 4096 128 3000 4 0 1 1 2>&1 | tail -3
echo "== sat single"; timeout 200 python bench/gpu_sat.py 4096 128 3000 4 0 1 0 2>&1 | tail -3
echo "== sat dual"; timeout 200 python bench/gpu_sat.py 4096 128 3000 4 0 1 1 2>&1 | tail -3
echo "== prof dual"; MPC_B200_LIB=mpc_ros_b200/lib/libmpc_b200_prof.so timeout 200 python bench/gpu_sat.py 4096 128 3000 4 0 1 1 2>&1 | tail -7
} > $O/r2j_dual.txt 2>&1
cat $O/r2j_dual.txt
