#!/bin/bash
# round 2, pass 2: the whole GPU tier without -x (which tests fail, not just the first)
O=gpurun_out; mkdir -p $O
timeout 2000 python -m pytest tests -m gpu -q 2>&1 | tail -80 > $O/r2b_pytest.log; tail -15 $O/r2b_pytest.log
