#!/bin/bash
# round 2, pass 19: queue pop issued at the end of P2 (its round trip overlaps the write-out and barrier B1b)
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -2 > $O/r2s_pytest.log; cat $O/r2s_pytest.log
{
echo "== sat dual"; timeout 200 python bench/gpu_sat.py 4096 128 3000 4 0 1 1 2>&1 | tail -2
echo "== sat single"; timeout 200 python bench/gpu_sat.py 4096 128 3000 4 0 1 0 2>&1 | tail -2
echo "== prof dual"; MPC_B200_LIB=mpc_ros_b200/lib/libmpc_b200_prof.so timeout 200 python bench/gpu_sat.py 4096 128 3000 4 0 1 1 2>&1 | tail -6
} > $O/r2s_sat.txt 2>&1
cat $O/r2s_sat.txt
