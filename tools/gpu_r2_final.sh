#!/bin/bash
# round 2, last 1-GPU visit: quick checks of the newest kernels (resize pair kernel with 32 / 64 planes per CTA, LC3D
# row kernel as the batch-8 default), then the round-end sequence the driver runs (GPU tests, smoke, reference arm,
# default bench) and a fresh launch list of the default bench command
mkdir -p gpurun_out
t() { python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], d['ms_per_step'], d['roofline']['frac'], d.get('clocks', {}).get('sm_mhz'))" "$1" "$2"; }
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "resize or lc3d" 2>&1 | tail -6 ) > gpurun_out/r2g_pytest_new.log 2>&1; tail -3 gpurun_out/r2g_pytest_new.log
for v in 32 64; do
  ( NRT_RESIZE_TZ=$v timeout 300 python bench.py --op resize --no-cpu-baseline ) > gpurun_out/r2g_op_resize_tz$v.json 2>> gpurun_out/r2g_bench.err
  t gpurun_out/r2g_op_resize_tz$v.json "resize pair TZ=$v"
done
( NRT_RESIZE_TZ=64 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "resize" 2>&1 | tail -3 ) > gpurun_out/r2g_pytest_resize_tz64.log 2>&1; tail -1 gpurun_out/r2g_pytest_resize_tz64.log
( timeout 300 python bench.py --op lc3d --lc-batch 8 --no-cpu-baseline ) > gpurun_out/r2g_op_lc3d_b8.json 2>> gpurun_out/r2g_bench.err
t gpurun_out/r2g_op_lc3d_b8.json "lc3d B=8 default"
( timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6 ) > gpurun_out/r2g_pytest_all.log 2>&1; tail -3 gpurun_out/r2g_pytest_all.log
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/r2g_smoke.log 2>&1; tail -1 gpurun_out/r2g_smoke.log
( timeout 300 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r2g_bench_reference.json 2>> gpurun_out/r2g_bench.err; cut -c1-300 gpurun_out/r2g_bench_reference.json
( timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2g_bench_default.json 2>> gpurun_out/r2g_bench.err; cut -c1-700 gpurun_out/r2g_bench_default.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2g_launches_bench.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2g_ncu_launch.log 2>&1
ls -la gpurun_out | grep r2g_ | wc -l
