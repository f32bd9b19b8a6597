#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 180 -x 2>&1 | tail -15 ) > gpurun_out/pytest_parity.log 2>&1
tail -5 gpurun_out/pytest_parity.log
for pack in 1 0 1 0; do
  for flow in iid smooth; do
    NRT_WARP_PACK=$pack timeout 200 python bench.py --steps 100 --warmup 10 --flow $flow --no-cpu-baseline --e2e-steps 1 > gpurun_out/b.json 2>/dev/null
    python -c "
import json; d=json.loads(open('gpurun_out/b.json').read().strip().splitlines()[-1]); print('pack=$pack flow=$flow ms/step %.4f frac %.3f' % (d['ms_per_step'], d['roofline']['frac']))"
  done
done
