#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -30 ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
for op in lc3d resize; do ( timeout 300 python bench.py --op $op --steps 10 --warmup 3 ) | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$op', d['ms_per_step'], d['roofline']['frac'])"; done
for b in 2 4 8; do ( timeout 300 python bench.py --op lc3d --lc-batch $b --steps 5 --warmup 3 ) | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('lc3d b$b', d['ms_per_step'], d['roofline']['frac'])"; done
( timeout 600 python tools/bench_misc.py ) 2>&1 | tail -12
