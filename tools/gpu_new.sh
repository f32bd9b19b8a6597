#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -40 ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log | cut -c1-200
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
for op in mi blur resize; do
  ( timeout 200 python bench.py --op $op --steps 30 --warmup 5 ) > gpurun_out/bench_$op.json 2> gpurun_out/bench_$op.err
  python -c "
import json; d=json.loads(open('gpurun_out/bench_$op.json').read().strip().splitlines()[-1]); print('$op ms/step %.4f frac %.3f' % (d['ms_per_step'], d['roofline']['frac']))"
done
for tz in 16 64; do
  ( NRT_RESIZE_TZ=$tz timeout 200 python bench.py --op resize --steps 30 --warmup 5 ) > gpurun_out/bench_resize_tz$tz.json 2>> gpurun_out/bench_resize.err
  python -c "
import json; d=json.loads(open('gpurun_out/bench_resize_tz$tz.json').read().strip().splitlines()[-1]); print('resize TZ=$tz ms/step %.4f frac %.3f' % (d['ms_per_step'], d['roofline']['frac']))"
done
( timeout 600 python tools/bench_new.py ) > gpurun_out/bench_new.txt 2>&1; grep -v "col mode [01]" gpurun_out/bench_new.txt
prof() {
  timeout 400 ncu --set full --clock-control none -k regex:$2 -s 3 -c 1 -o gpurun_out/prof_$1 -f \
    python bench.py $3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$1.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_$1.ncu-rep gpurun_out/ncu_summary_$1.txt > /dev/null 2>&1
}
prof mi mi_hist_mma "--op mi"
prof blur_row sepconv_row "--op blur"
prof resize resize3d "--op resize"
( timeout 200 python bench.py --steps 50 --warmup 5 ) > gpurun_out/bench_warp.json 2> gpurun_out/bench_warp.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_warp.json').read().strip().splitlines()[-1]); print('warp ms/step %.4f frac %.3f e2e %.3e' % (d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']))"
