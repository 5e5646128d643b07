#!/bin/bash
# first GPU pass of round 2: tests, smoke, the redefined bench line, phase-cycle profile of the solve kernel, config 3 one-shot
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader; nproc
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $O/r2a_pytest.log; tail -5 $O/r2a_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > $O/r2a_bench.json 2> $O/r2a_bench.err; tail -3 $O/r2a_bench.err; cut -c1-400 $O/r2a_bench.json
MPC_B200_LIB=mpc_ros_b200/lib/libmpc_b200_prof.so timeout 300 python bench/gpu_sat.py 4096 128 3000 4 > $O/r2a_phase_cycles.txt 2>&1; cat $O/r2a_phase_cycles.txt
timeout 300 mpc_ros_b200/lib/mpc_bench multi 1 65536 5 | tail -1 > $O/r2a_config3_1gpu.json; cat $O/r2a_config3_1gpu.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2a_bench_reference.json 2>/dev/null; cut -c1-200 $O/r2a_bench_reference.json
