#!/bin/bash
# round 2: C = 1 tile kernel with the box-following decision handed over through an mbarrier (no block barrier)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "warp or interpn or march or transform" 2>&1 | tail -6 ) > gpurun_out/r2j_pytest_warp.log 2>&1; tail -3 gpurun_out/r2j_pytest_warp.log
( SWEEP_ONLY=c1 timeout 600 python tools/sweep_r2.py 2>&1 ) > gpurun_out/r2j_sweep_c1.txt 2>&1; cat gpurun_out/r2j_sweep_c1.txt
( SWEEP_ONLY=c1 timeout 600 python tools/sweep_r2.py 2>&1 | grep "iid3\|smooth3" | grep linear ) > gpurun_out/r2j_sweep_c1_rep.txt 2>&1; cat gpurun_out/r2j_sweep_c1_rep.txt
