#!/bin/bash
# round 2: CCE kernel with compile-time lanes per row and 2 / 4 rows per thread against the run-time-q kernel
mkdir -p gpurun_out
t() { python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], d['ms_per_step'], d['roofline']['frac'], d.get('clocks', {}).get('sm_mhz'))" "$1" "$2"; }
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_grad.py -m gpu -q --timeout 600 -k "cce or resize" 2>&1 | tail -8 ) > gpurun_out/r2m_pytest_cce.log 2>&1; tail -4 gpurun_out/r2m_pytest_cce.log
for rep in 1 2; do
 for v in 0 2 4; do
  ( NRT_CCE_UNROLL=$v timeout 300 python bench.py --op cce --no-cpu-baseline ) > gpurun_out/r2m_op_cce_u${v}_${rep}.json 2>> gpurun_out/r2m_bench.err
  t gpurun_out/r2m_op_cce_u${v}_${rep}.json "cce unroll=$v rep $rep"
 done
done
