#!/bin/bash
# round 2, visit 4 (1 GPU): resize tile kernel, LC3D variants
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "resize or lc3d" 2>&1 | tail -15 ) > gpurun_out/r2v4_pytest.log 2>&1; tail -4 gpurun_out/r2v4_pytest.log
one() { ( env $1 timeout 300 python bench.py --op $2 --no-cpu-baseline ) > gpurun_out/r2v4_tmp.json 2>> gpurun_out/r2v4.err; python -c "
import json; d=json.loads(open('gpurun_out/r2v4_tmp.json').read().strip().splitlines()[-1]); print('%-40s %-28s ms %.4f frac %.3f' % ('$1', '$2', d['ms_per_step'], d['roofline']['frac']))"; }
one "NRT_RESIZE_TILE=1" "resize"
one "NRT_RESIZE_TILE=0" "resize"
one "NRT_RESIZE_TZ=16" "resize"
one "NRT_RESIZE_TZ=64" "resize"
one "NRT_LC3D_PATCH1=1" "lc3d"
one "NRT_LC3D_PATCH1=0" "lc3d"
one "NRT_LC3D_B8=24" "lc3d --lc-batch 8"
one "NRT_LC3D_B8=42" "lc3d --lc-batch 8"
one "NRT_LC3D_B8=24 NRT_LC3D_WARPS=3" "lc3d --lc-batch 8"
one "NRT_LC3D_B8=24 NRT_LC3D_WARPS=2" "lc3d --lc-batch 8"
one "X=1" "lc3d --lc-batch 4"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:resize3d -s 2 -c 1 -o gpurun_out/r2v4_prof_resize -f python bench.py --op resize --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2v4_ncu_resize.log 2>&1
tail -3 gpurun_out/r2v4.err
