#!/bin/bash
# round 2, pass 13: two-step reciprocal + reduction of the partial sums on both half-warps -- tests, saturated A/B, phases
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 > $O/r2m_pytest.log; cat $O/r2m_pytest.log
{
echo "== sat dual"; timeout 200 python bench/gpu_sat.py 4096 128 3000 4 0 1 1 2>&1 | tail -2
echo "== sat single"; timeout 200 python bench/gpu_sat.py 4096 128 3000 4 0 1 0 2>&1 | tail -2
echo "== prof dual"; MPC_B200_LIB=mpc_ros_b200/lib/libmpc_b200_prof.so timeout 200 python bench/gpu_sat.py 4096 128 3000 4 0 1 1 2>&1 | tail -6
echo "== prof single"; MPC_B200_LIB=mpc_ros_b200/lib/libmpc_b200_prof.so timeout 200 python bench/gpu_sat.py 4096 128 3000 4 0 1 0 2>&1 | tail -6
} > $O/r2m_sat.txt 2>&1
cat $O/r2m_sat.txt
