#!/bin/bash
# round 2, pass 4: launch audit (every launch writes its own status array) and A/B of the library before / after the SOC
O=gpurun_out; mkdir -p $O
for lib in libmpc_b200_pre.so libmpc_b200.so; do
  for rep in 1 2; do
    echo "== $lib"; MPC_B200_LIB=mpc_ros_b200/lib/$lib timeout 300 python bench/gpu_sat.py 4096 128 3000 4 2>&1 | tail -3
  done
done > $O/r2d_ab.txt 2>&1
cat $O/r2d_ab.txt
