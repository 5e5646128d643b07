#!/bin/bash
# round 2, pass 20: feasibility restoration in the kernel -- tests, saturated throughput, config-3 batch
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 > $O/r2t_pytest.log; cat $O/r2t_pytest.log
{
timeout 120 python bench/dual_check.py 20000 8 2>&1 | tail -9
echo "== sat dual"; timeout 200 python bench/gpu_sat.py 4096 128 3000 4 0 1 1 2>&1 | tail -2
echo "== sat single"; timeout 200 python bench/gpu_sat.py 4096 128 3000 4 0 1 0 2>&1 | tail -2
echo "== config 3 one-shot"; timeout 300 mpc_ros_b200/lib/mpc_bench multi 1 65536 7 | tail -1 | cut -c1-330
} > $O/r2t_resto.txt 2>&1
cat $O/r2t_resto.txt
