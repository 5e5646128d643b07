#!/bin/bash
# round 2, pass 16: both control warps of the dual-group kernel on sub-partition 0 (warps 0 and 4)
O=gpurun_out; mkdir -p $O
{
timeout 120 python bench/dual_check.py 20000 8 2>&1 | tail -9
timeout 600 python -m pytest tests -m gpu -q -x -k "dual or special or sweep_config3 or warm or full_size" 2>&1 | tail -2
echo "== sat dual"; timeout 200 python bench/gpu_sat.py 4096 128 3000 4 0 1 1 2>&1 | tail -2
echo "== sat dual"; timeout 200 python bench/gpu_sat.py 4096 128 3000 4 0 1 1 2>&1 | tail -2
echo "== prof dual"; MPC_B200_LIB=mpc_ros_b200/lib/libmpc_b200_prof.so timeout 200 python bench/gpu_sat.py 4096 128 3000 4 0 1 1 2>&1 | tail -6
} > $O/r2p_ctrl04.txt 2>&1
cat $O/r2p_ctrl04.txt
