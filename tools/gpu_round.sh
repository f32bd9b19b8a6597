#!/bin/bash
# One gpurun visit on a 1-GPU B200 box.  Everything lands in gpurun_out/ (merged back).
#   MODE=check   (default) tests with per-test timeout, smoke, bench lines
#   MODE=profile + ncu launch list and one `--set full` capture per hot kernel (for profiles/)
#   MODE=sweep   + tools/sweep_warp.py (tile configs x flows) and tools/bench_misc.py
MODE=${MODE:-check}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -40 ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
( timeout 300 python bench.py --steps 50 --warmup 5 ) > gpurun_out/bench_warp.json 2> gpurun_out/bench_warp.err; cut -c1-700 gpurun_out/bench_warp.json
( timeout 200 python bench.py --steps 50 --warmup 5 --flow smooth --no-cpu-baseline ) > gpurun_out/bench_warp_smooth.json 2>> gpurun_out/bench_warp.err
( timeout 200 python bench.py --steps 50 --warmup 5 --method nearest --no-cpu-baseline ) > gpurun_out/bench_warp_nearest.json 2>> gpurun_out/bench_warp.err
for op in dice cce lc3d resize mi mi_segs blur; do
  ( timeout 300 python bench.py --op $op --steps 20 --warmup 3 ) > gpurun_out/bench_$op.json 2> gpurun_out/bench_$op.err
done
( timeout 300 python bench.py --op lc3d --lc-batch 8 --steps 5 --warmup 3 ) > gpurun_out/bench_lc3d_b8.json 2>> gpurun_out/bench_lc3d.err
( timeout 200 python bench.py --impl reference --steps 10 --warmup 2 ) > gpurun_out/bench_reference.json 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get('roofline') or {}
        print('%-34s ms/step %8.4f  value %.4e %s  frac %s  e2e %s' % (f.split('/')[-1], d['ms_per_step'], d['value'], d['unit'],
              ('%.3f' % r['frac']) if r else '-', (d.get('e2e') or {}).get('value')))
    except Exception as e:
        print(f, 'unreadable:', e)
PY
if [ "$MODE" = "sweep" ]; then
  ( timeout 600 python tools/sweep_warp.py ) > gpurun_out/sweep.txt 2>&1; grep -E "linear" gpurun_out/sweep.txt
  ( timeout 600 python tools/bench_misc.py ) > gpurun_out/bench_misc.txt 2>&1; cat gpurun_out/bench_misc.txt
fi
if [ "$MODE" = "profile" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_warp.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch.log 2>&1
  prof() {  # name, kernel regex, bench args
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -o gpurun_out/prof_$1 -f \
      python bench.py $3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$1.log 2>&1
  }
  prof warp warp3d_tile "--e2e-steps 1"
  prof dice dice_sums "--op dice"
  prof cce cce_vec4 "--op cce"
  prof lc3d lc3d_stream "--op lc3d"
  prof lc3d_b8 lc3d_stream "--op lc3d --lc-batch 8"
  prof resize resize3d "--op resize"
  ls -la gpurun_out | grep -E "prof|launches"
fi
