#!/bin/bash
# profiles/build_prof.sh -- the NMPC_PROFILE build of the library (per-phase clock64 counters in the solve kernel),
# used only by bench/gpu_sat.py / bench/gpu_prof.py through MPC_B200_LIB=mpc_ros_b200/lib/libmpc_b200_prof.so.
cd "$(dirname "$0")/.." || exit 1
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DNMPC_PROFILE -shared \
    -o mpc_ros_b200/lib/libmpc_b200_prof.so mpc_ros_b200/csrc/mpc_b200.cu mpc_ros_b200/csrc/params.cpp
