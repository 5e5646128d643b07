#!/bin/bash
# profiles/r2_evidence.sh -- round-2 evidence in one gpurun call (1 GPU):
#   gpurun --timeout 2400 -- 'bash profiles/r2_evidence.sh'
# (ncu passes run only after the same command has exited 0 without ncu; numbers printed under ncu are not used.)
T=r2f; O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader; nproc
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 > $O/${T}_pytest.log; cat $O/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/${T}_smoke.txt
timeout 900 python bench.py > $O/${T}_bench_default.json 2> $O/${T}_bench_default.err; cut -c1-200 $O/${T}_bench_default.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > $O/${T}_bench_reference.json 2>/dev/null
CMD="python bench.py --steps 1 --warmup 3 --batches-per-step 16 --streams 8 --no-extras --no-cpu-baseline --e2e-inflight 4"
$CMD > $O/${T}_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $O/${T}_launches.csv $CMD > $O/${T}_ncu_list.log 2>&1
$CMD > $O/${T}_plainb.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:nmpc_solve -s 10 -c 1 \
    -f -o $O/${T}_prof_solve $CMD > $O/${T}_ncu_full.log 2>&1
tail -1 $O/${T}_ncu_full.log
ncu -i $O/${T}_prof_solve.ncu-rep --page raw --csv > $O/${T}_solve_kernel_ncu_raw.csv 2>/dev/null
MPC_B200_LIB=mpc_ros_b200/lib/libmpc_b200_prof.so timeout 300 python bench/gpu_sat.py 4096 128 3000 4 > $O/${T}_phase_cycles.txt 2>&1
timeout 300 python bench/gpu_sat.py 4096 128 3000 4 > $O/${T}_gpu_sat.txt 2>&1
MPC_B200_LIB=mpc_ros_b200/lib/libmpc_b200_prof.so timeout 300 python bench/config4_prof.py 16384 100 > $O/${T}_config4_prof.json 2>&1
timeout 120 mpc_ros_b200/lib/mpc_bench latency 10000 2>/dev/null | tail -1 > $O/${T}_config1_latency.json
timeout 300 python bench/config4.py 16384 100 > $O/${T}_config4.json 2>&1
timeout 900 python bench/closed_loop.py 1024 500 --oracle-subset 32 > $O/${T}_config5.json 2>$O/${T}_config5.err
timeout 600 python bench/closed_loop.py 1024 500 --device > $O/${T}_config5_device.json 2>>$O/${T}_config5.err
timeout 900 python tests/parity_sweep.py 16384 > $O/${T}_parity_sweep.json 2> $O/${T}_parity_sweep.err
ls -la $O | tail -30
