#!/bin/bash
# round 2, pass 11: dual-group kernel as the default -- GPU test tier, saturated A/B, bench line
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > $O/r2k_pytest.log; cat $O/r2k_pytest.log
{
echo "== sat dual"; timeout 200 python bench/gpu_sat.py 4096 128 3000 4 0 1 1 2>&1 | tail -2
echo "== sat single"; timeout 200 python bench/gpu_sat.py 4096 128 3000 4 0 1 0 2>&1 | tail -2
echo "== prof dual"; MPC_B200_LIB=mpc_ros_b200/lib/libmpc_b200_prof.so timeout 200 python bench/gpu_sat.py 4096 128 3000 4 0 1 1 2>&1 | tail -6
} > $O/r2k_dual.txt 2>&1
cat $O/r2k_dual.txt
timeout 600 python bench.py --no-cpu-baseline > $O/r2k_bench.json 2> $O/r2k_bench.err; cut -c1-600 $O/r2k_bench.json; tail -3 $O/r2k_bench.err
