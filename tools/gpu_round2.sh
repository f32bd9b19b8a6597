#!/bin/bash
# focused visit: tests (per-test timeout), quick kernel sweep, ncu capture of the warp kernel
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --timeout 120 2>&1 | tail -150 ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
( timeout 120 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
( SWEEP_QUICK=1 timeout 300 python tools/sweep_warp.py ) > gpurun_out/sweep_quick.txt 2>&1; cat gpurun_out/sweep_quick.txt
for op in lc3d; do ( timeout 300 python bench.py --op $op --steps 10 --warmup 3 ) | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$op', d['ms_per_step'], d['roofline']['frac'])"; done
( timeout 300 python bench.py --op lc3d --lc-batch 8 --steps 5 --warmup 3 ) | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('lc3d b8', d['ms_per_step'], d['roofline']['frac'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:warp3d_tile -s 3 -c 1 -o gpurun_out/prof_warp -f \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_warp.log 2>&1
ls -la gpurun_out | grep prof
