#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_mi_conv.py tests/test_gpu_parity.py -m gpu -q --timeout 180 2>&1 | tail -60 ) > gpurun_out/pytest_new.log 2>&1
tail -40 gpurun_out/pytest_new.log | cut -c1-220
timeout 200 python bench.py --op resize --steps 30 --warmup 5 > gpurun_out/bench_resize.json 2> gpurun_out/bench_resize.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_resize.json').read().strip().splitlines()[-1]); print('resize ms/step %.4f frac %.3f' % (d['ms_per_step'], d['roofline']['frac']))"
timeout 200 python bench.py --op blur --steps 30 --warmup 5 > gpurun_out/bench_blur.json 2> gpurun_out/bench_blur.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_blur.json').read().strip().splitlines()[-1]); print('blur ms/step %.4f frac %.3f' % (d['ms_per_step'], d['roofline']['frac']))"
