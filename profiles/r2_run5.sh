#!/bin/bash
# round 2, pass 5: where the N = 100 cycle goes (profile build)
O=gpurun_out; mkdir -p $O
MPC_B200_LIB=mpc_ros_b200/lib/libmpc_b200_prof.so timeout 300 python bench/config4_prof.py 16384 100 > $O/r2e_config4_prof.json 2>&1; cat $O/r2e_config4_prof.json
timeout 300 python bench/config4.py > $O/r2e_config4.json 2>&1; cat $O/r2e_config4.json
