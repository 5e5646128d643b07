#!/bin/bash
# round 2, pass 14: one stage per stage thread in narrow CTAs (latency mode) -- tests, config-1 latency, mid-size batches
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 > $O/r2n_pytest.log; cat $O/r2n_pytest.log
{
timeout 120 mpc_ros_b200/lib/mpc_bench latency 10000 2>/dev/null | tail -1
for b in 64 512 1200 2048; do timeout 120 mpc_ros_b200/lib/mpc_bench batch $b 200 2>/dev/null | tail -1 | cut -c1-300; MPC_BENCH_TWO_STAGES=1 timeout 120 mpc_ros_b200/lib/mpc_bench batch $b 200 2>/dev/null | tail -1 | cut -c1-300; done
} > $O/r2n_latency.txt 2>&1
cat $O/r2n_latency.txt
