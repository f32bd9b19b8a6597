#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6 ) > gpurun_out/r2v5_pytest_all.log 2>&1; tail -3 gpurun_out/r2v5_pytest_all.log
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/r2v5_smoke.log 2>&1; tail -1 gpurun_out/r2v5_smoke.log
( SWEEP_ONLY=c1 timeout 600 python tools/sweep_r2.py ) > gpurun_out/r2v5_sweep_c1.txt 2>&1; grep linear gpurun_out/r2v5_sweep_c1.txt
one() { ( env $1 timeout 300 python bench.py --op $2 --no-cpu-baseline ) > gpurun_out/r2v5_tmp.json 2>> gpurun_out/r2v5.err; python -c "
import json; d=json.loads(open('gpurun_out/r2v5_tmp.json').read().strip().splitlines()[-1]); print('%-28s %-28s ms %.4f frac %.3f' % ('$1', '$2', d['ms_per_step'], d['roofline']['frac']))"; }
one "NRT_RESIZE_MINB=3" "resize"
one "NRT_RESIZE_MINB=1" "resize"
one "NRT_MARCH_QPT=2" "warp_mc --channels 16"
one "NRT_MARCH_QPT=1" "warp_mc --channels 16"
one "X=1" "warp_mc --channels 16 --flow smooth"
one "X=1" "warp_mc --channels 3"
one "X=1" "warp_mc --channels 4"
tail -3 gpurun_out/r2v5.err
