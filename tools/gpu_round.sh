#!/bin/bash
# One gpurun visit: tests, smoke, kernel sweep, bench lines, ncu launch list + full capture.
# Everything is logged under gpurun_out/ (merged back by gpurun).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -40 ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
( timeout 600 python tools/sweep_warp.py ) > gpurun_out/sweep.txt 2>&1; tail -60 gpurun_out/sweep.txt
( timeout 300 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_warp.json 2> gpurun_out/bench_warp.err; cat gpurun_out/bench_warp.json
( timeout 200 python bench.py --steps 20 --warmup 5 --flow smooth --halo 4 --no-cpu-baseline ) > gpurun_out/bench_warp_smooth.json 2>> gpurun_out/bench_warp.err
for op in dice cce lc3d resize; do
  ( timeout 300 python bench.py --op $op --steps 10 --warmup 3 ) > gpurun_out/bench_$op.json 2> gpurun_out/bench_$op.err; cat gpurun_out/bench_$op.json
done
( timeout 300 python bench.py --op lc3d --lc-batch 8 --steps 5 --warmup 3 ) > gpurun_out/bench_lc3d_b8.json 2>> gpurun_out/bench_lc3d.err; cat gpurun_out/bench_lc3d_b8.json
( timeout 200 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/bench_reference.json 2>&1; cat gpurun_out/bench_reference.json
for w in 2 6 7; do echo "lc3d warps=$w"; ( NRT_LC3D_WARPS=$w timeout 300 python bench.py --op lc3d --steps 10 --warmup 3 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'])" ); done
for w in 2 6 7; do echo "lc3d b8 warps=$w"; ( NRT_LC3D_WARPS=$w timeout 300 python bench.py --op lc3d --lc-batch 8 --steps 5 --warmup 3 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'])" ); done
if [ "${NCU:-1}" = "0" ]; then ls gpurun_out | wc -l; exit 0; fi
# ncu: launch list of the bench command, then one full capture of the top kernel of each op
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_warp.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:warp3d_tile -s 3 -c 2 -o gpurun_out/prof_warp -f \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_warp.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dice_sums -s 3 -c 1 -o gpurun_out/prof_dice -f \
  python bench.py --op dice --steps 2 --warmup 3 > gpurun_out/ncu_full_dice.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lc3d_stream -s 3 -c 1 -o gpurun_out/prof_lc3d -f \
  python bench.py --op lc3d --steps 2 --warmup 3 > gpurun_out/ncu_full_lc3d.log 2>&1
ls -la gpurun_out | tail -30
