#!/bin/bash
# round 2, visit 2 (1 GPU): debug the LC3D patch kernel fault, re-check everything, sweep the new variants, full bench line
mkdir -p gpurun_out
( timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 250 -k "lc3d_shared_weights" 2>&1 | tail -60 ) > gpurun_out/r2v2_sanitize_lc3d.log 2>&1
grep -m3 -E "Invalid|Error|error|passed|failed" gpurun_out/r2v2_sanitize_lc3d.log
( timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -60 ) > gpurun_out/r2v2_pytest_all.log 2>&1
tail -25 gpurun_out/r2v2_pytest_all.log
( SWEEP_ONLY=quick timeout 600 python tools/sweep_r2.py ) > gpurun_out/r2v2_sweep.txt 2>&1; cat gpurun_out/r2v2_sweep.txt | grep -v nearest
for x2 in 1 0; do ( NRT_RESIZE_X2=$x2 timeout 200 python bench.py --op resize --no-cpu-baseline ) > gpurun_out/r2v2_resize_x2_$x2.json 2>> gpurun_out/r2v2_bench.err; done
for pt in 1 0; do ( NRT_LC3D_PATCH=$pt timeout 300 python bench.py --op lc3d --lc-batch 8 --no-cpu-baseline ) > gpurun_out/r2v2_lc3d_b8_patch_$pt.json 2>> gpurun_out/r2v2_bench.err; done
( NRT_LC3D_PATCH=1 timeout 300 python bench.py --op lc3d --lc-batch 2 --no-cpu-baseline ) > gpurun_out/r2v2_lc3d_b2_patch_1.json 2>> gpurun_out/r2v2_bench.err
( timeout 600 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2v2_bench_default.json 2>> gpurun_out/r2v2_bench.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2v2_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get('roofline') or {}
        print('%-40s ms/step %8.4f  frac %s' % (f.split('/')[-1], d['ms_per_step'], ('%.3f' % r['frac']) if r else '-'))
        if 'ops' in d:
            for k, v in d['ops'].items():
                print('   ops.%-10s %s' % (k, ('ms %.4f frac %.3f cpu %s' % (v['ms_per_step'], v['roofline']['frac'], (v.get('cpu_baseline') or {}).get('value'))) if 'error' not in v else v['error']))
            print('   long_run', d.get('long_run'), 'e2e', d['e2e']['value'], 'cpu', (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e:
        print(f, 'unreadable:', e)
PY
tail -5 gpurun_out/r2v2_bench.err
