#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -40 ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
( timeout 120 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
( timeout 600 python tools/sweep_warp.py ) > gpurun_out/sweep.txt 2>&1; grep -E "linear" gpurun_out/sweep.txt
( timeout 200 python bench.py --steps 50 --warmup 5 --flow smooth --no-cpu-baseline ) > gpurun_out/bench_warp_smooth.json 2> gpurun_out/bench_warp.err; cut -c1-400 gpurun_out/bench_warp_smooth.json
