#!/bin/bash
# One long gpurun visit: full GPU test suite, smoke, every bench line, dev sweeps, and the ncu
# evidence for profiles/ (launch list of the headline bench + one --set full capture per hot kernel).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -40 ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
( timeout 300 python bench.py --steps 50 --warmup 5 ) > gpurun_out/bench_warp.json 2> gpurun_out/bench_warp.err
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
# ---- ncu evidence
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_warp.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch.log 2>&1
prof() {  # name, kernel regex, bench args
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -o gpurun_out/prof_$1 -f \
    python bench.py $3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$1.log 2>&1
}
prof warp warp3d_tile "--e2e-steps 1"
prof resize resize3d "--op resize"
prof mi mi_hist_mma "--op mi"
prof mi_segs mi_hist_mma "--op mi_segs"
prof blur_col sepconv_col4 "--op blur"
prof blur_row sepconv_row "--op blur"
# summaries are made on the box (ncu -i needs no GPU); big reports other than the headline kernel's are dropped
# so that the visit stays inside the 64 MiB that gpurun merges back
for n in warp dice cce lc3d lc3d_b8 resize mi mi_segs blur_col blur_row; do
  [ -f gpurun_out/prof_$n.ncu-rep ] && python tools/ncu_summary.py gpurun_out/prof_$n.ncu-rep gpurun_out/ncu_summary_$n.txt > /dev/null 2>&1
done
for n in resize mi mi_segs blur_col blur_row lc3d lc3d_b8; do
  f=gpurun_out/prof_$n.ncu-rep
  [ -f $f ] && [ $(stat -c %s $f) -gt 6000000 ] && rm -f $f
done
rm -f gpurun_out/b.json
ls -la gpurun_out | grep -E "prof|launches|summary"
du -sh gpurun_out

# ---- dev sweeps (after the evidence is safe)
for tz in 8 16 32; do
  ( NRT_RESIZE_TZ=$tz timeout 200 python bench.py --op resize --steps 30 --warmup 5 ) > gpurun_out/bench_resize_tz$tz.json 2>> gpurun_out/bench_resize.err
done
for tz in 8 16 32; do python -c "
import json; d=json.loads(open('gpurun_out/bench_resize_tz$tz.json').read().strip().splitlines()[-1]); print('resize TZ=$tz ms/step %.4f frac %.3f' % (d['ms_per_step'], d['roofline']['frac']))"; done
( timeout 600 python tools/bench_new.py ) > gpurun_out/bench_new.txt 2>&1; cat gpurun_out/bench_new.txt

# lowest priority: memcheck over the newest kernels (small cases only)
( timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_mi_conv.py -m gpu -q -x --timeout 200 \
    -k "golden or errors or gradient or subsample" 2>&1 | tail -15 ) > gpurun_out/sanitizer_new.txt 2>&1
tail -4 gpurun_out/sanitizer_new.txt
