#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_mi_conv.py tests/test_gpu_parity.py -m gpu -q --timeout 180 2>&1 | tail -30 ) > gpurun_out/pytest_new.log 2>&1
tail -8 gpurun_out/pytest_new.log
( SKIP_MI=1 timeout 600 python tools/bench_new.py ) > gpurun_out/bench_new.txt 2>&1; cat gpurun_out/bench_new.txt
for op in resize blur; do
  ( timeout 200 python bench.py --op $op --steps 20 --warmup 3 ) > gpurun_out/bench_$op.json 2> gpurun_out/bench_$op.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_$op.json').read().strip().splitlines()[-1])
    print('$op', 'ms/step %.4f' % d['ms_per_step'], 'value %.4e' % d['value'], 'frac %.3f' % d['roofline']['frac'])
except Exception as e:
    print('$op unreadable', e); print(open('gpurun_out/bench_$op.err').read()[-1500:])
PY
done
