#!/bin/bash
# round 2, pass 7: four-lane split sweeps for CTAs of <= 8 lanes -- GPU tier, config 4 (N = 100) profile + timing, config 1 latency
O=gpurun_out; mkdir -p $O
timeout 2000 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/r2g_pytest.log; tail -6 $O/r2g_pytest.log
MPC_B200_LIB=mpc_ros_b200/lib/libmpc_b200_prof.so timeout 300 python bench/config4_prof.py 16384 100 > $O/r2g_config4_prof.json 2>&1; cat $O/r2g_config4_prof.json
timeout 300 python bench/config4.py > $O/r2g_config4.json 2>&1; cat $O/r2g_config4.json
timeout 120 mpc_ros_b200/lib/mpc_bench latency 3000 2>&1 | tail -2
