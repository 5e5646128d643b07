#!/bin/bash
# round 2, pass 6: 4x4 Riccati sweeps (etheta eliminated) -- saturated throughput A/B and phase cycles
O=gpurun_out; mkdir -p $O
for lib in libmpc_b200_pre.so libmpc_b200.so; do
  echo "== $lib"; MPC_B200_LIB=mpc_ros_b200/lib/$lib timeout 300 python bench/gpu_sat.py 4096 128 3000 4 2>&1 | tail -2
done > $O/r2f_ab.txt 2>&1
MPC_B200_LIB=mpc_ros_b200/lib/libmpc_b200_prof.so timeout 300 python bench/gpu_sat.py 4096 128 3000 4 >> $O/r2f_ab.txt 2>&1
MPC_B200_LIB=mpc_ros_b200/lib/libmpc_b200_prof.so timeout 300 python bench/config4_prof.py 16384 100 >> $O/r2f_ab.txt 2>&1
cat $O/r2f_ab.txt
