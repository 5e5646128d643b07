#!/bin/bash
# round 2, pass 25: two-part evaluation in the dual-group kernel (part A runs while the control thread does the adjoint sweep)
O=gpurun_out; mkdir -p $O
{
timeout 120 python bench/dual_check.py 20000 8 2>&1 | tail -9
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for d in 1 0 1; do timeout 200 python bench/gpu_sat.py 4096 128 3000 4 0 1 $d 2>&1 | tail -1; done
MPC_B200_LIB=mpc_ros_b200/lib/libmpc_b200_prof.so timeout 200 python bench/gpu_sat.py 4096 128 3000 4 0 1 1 2>&1 | tail -4 | head -3
} > $O/r2z_split.txt 2>&1
cat $O/r2z_split.txt
