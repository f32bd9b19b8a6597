#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "march or slabs or batch_and_channels" 2>&1 | tail -8 ) > gpurun_out/r2v9_pytest.log 2>&1; tail -3 gpurun_out/r2v9_pytest.log
one() { ( env $1 timeout 300 python bench.py --op $2 --no-cpu-baseline ) > gpurun_out/r2v9_tmp.json 2>> gpurun_out/r2v9.err; python -c "
import json; d=json.loads(open('gpurun_out/r2v9_tmp.json').read().strip().splitlines()[-1]); print('%-22s %-40s ms %.4f frac %.3f' % ('$1', '$2', d['ms_per_step'], d['roofline']['frac']))"; }
one "NRT_MARCH_GROUPS=2" "warp_mc --channels 16"
one "NRT_MARCH_GROUPS=1" "warp_mc --channels 16"
one "NRT_MARCH_GROUPS=2" "warp_mc --channels 16 --flow smooth"
one "NRT_MARCH_GROUPS=2" "warp_mc --channels 8"
one "NRT_MARCH_GROUPS=2" "warp_mc --channels 8 --flow smooth"
one "NRT_MARCH_GROUPS=2" "warp_mc --channels 32"
tail -3 gpurun_out/r2v9.err
