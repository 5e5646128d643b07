#!/bin/bash
# round 2, pass 3: second-order correction in the kernel -- GPU tier, config 3 one-shot, bench line
O=gpurun_out; mkdir -p $O
timeout 2000 python -m pytest tests -m gpu -q 2>&1 | tail -60 > $O/r2c_pytest.log; tail -8 $O/r2c_pytest.log
timeout 300 mpc_ros_b200/lib/mpc_bench multi 1 65536 5 | tail -1 > $O/r2c_config3_1gpu.json; cat $O/r2c_config3_1gpu.json
timeout 600 python bench.py > $O/r2c_bench.json 2> $O/r2c_bench.err; tail -3 $O/r2c_bench.err; cut -c1-300 $O/r2c_bench.json
