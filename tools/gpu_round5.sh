#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -30 ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
( timeout 300 python bench.py --steps 50 --warmup 5 ) > gpurun_out/bench_warp.json 2> gpurun_out/bench_warp.err; cat gpurun_out/bench_warp.json | cut -c1-2500
( timeout 200 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/bench_reference.json 2>&1; cut -c1-300 gpurun_out/bench_reference.json
