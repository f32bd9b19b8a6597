#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "lc3d" 2>&1 | tail -5 ) > gpurun_out/r2v6_pytest.log 2>&1; tail -2 gpurun_out/r2v6_pytest.log
one() { ( env $1 timeout 300 python bench.py --op $2 --no-cpu-baseline ) > gpurun_out/r2v6_tmp.json 2>> gpurun_out/r2v6.err; python -c "
import json; d=json.loads(open('gpurun_out/r2v6_tmp.json').read().strip().splitlines()[-1]); print('%-36s %-24s ms %.4f frac %.3f' % ('$1', '$2', d['ms_per_step'], d['roofline']['frac']))"; }
one "NRT_LC3D_B8=22" "lc3d --lc-batch 8"
one "NRT_LC3D_B8=24" "lc3d --lc-batch 8"
one "NRT_LC3D_B8=22 NRT_LC3D_WARPS=4" "lc3d --lc-batch 8"
one "NRT_LC3D_B8=22 NRT_LC3D_WARPS=3" "lc3d --lc-batch 8"
timeout 300 ncu --set full --clock-control none -k regex:lc3d_patch -s 2 -c 1 -o gpurun_out/r2v6_prof_lc3d_pair -f env NRT_LC3D_B8=22 python bench.py --op lc3d --lc-batch 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2v6_ncu.log 2>&1
tail -3 gpurun_out/r2v6.err
