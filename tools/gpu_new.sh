#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -60 ) > gpurun_out/pytest_gpu.log 2>&1
tail -30 gpurun_out/pytest_gpu.log | cut -c1-220
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
b() {  # label, env, args
  ( env $2 timeout 200 python bench.py $3 ) > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/bench_tmp.json').read().strip().splitlines()[-1]); print('$1 ms/step %.4f frac %.3f' % (d['ms_per_step'], d['roofline']['frac']))
except Exception as e:
    print('$1 failed', e); print(open('gpurun_out/bench_tmp.err').read()[-800:])"
}
b "blur fused s1" "NRT_BLUR_FUSED=1" "--op blur --steps 30 --warmup 5"
b "blur passes s1" "NRT_BLUR_FUSED=0" "--op blur --steps 30 --warmup 5"
b "blur fused s2" "NRT_BLUR_FUSED=1" "--op blur --sigma 2 --steps 30 --warmup 5"
b "blur passes s2" "NRT_BLUR_FUSED=0" "--op blur --sigma 2 --steps 30 --warmup 5"
b "lc3d b8 packed" "NRT_LC3D_FFMA2=1" "--op lc3d --lc-batch 8 --steps 5 --warmup 3"
b "lc3d b8 scalar" "NRT_LC3D_FFMA2=0" "--op lc3d --lc-batch 8 --steps 5 --warmup 3"
b "lc3d b2 packed" "NRT_LC3D_FFMA2=1" "--op lc3d --lc-batch 2 --steps 5 --warmup 3"
b "lc3d b2 scalar" "NRT_LC3D_FFMA2=0" "--op lc3d --lc-batch 2 --steps 5 --warmup 3"
b "mi" "A=1" "--op mi --steps 30 --warmup 5"
b "resize" "A=1" "--op resize --steps 30 --warmup 5"
b "warp" "A=1" "--steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 2"
b "warp" "A=1" "--steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 2"
