#!/bin/bash
# round 2, pass 15: warm latency-mode kernels in the closed loop (BASELINE config 5)
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 > $O/r2o_pytest.log; cat $O/r2o_pytest.log
timeout 900 python bench/closed_loop.py 1024 500 --oracle-subset 32 > $O/r2o_config5.json 2>$O/r2o_config5.err
timeout 600 python bench/closed_loop.py 1024 500 --device > $O/r2o_config5_device.json 2>>$O/r2o_config5.err
cut -c1-400 $O/r2o_config5.json; cat $O/r2o_config5_device.json; tail -3 $O/r2o_config5.err
