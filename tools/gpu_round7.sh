#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -30 ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
( timeout 600 python tools/bench_misc.py ) 2>&1 | grep -E "warp|dice|cce"
( SWEEP_QUICK=1 timeout 300 python tools/sweep_warp.py ) 2>&1 | grep -E "linear  tile cfg2 halo3"
