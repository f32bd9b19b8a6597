#!/bin/bash
# official-style run: bench lines for every op, launch list + full ncu captures for profiles/
mkdir -p gpurun_out
( timeout 300 python bench.py --steps 50 --warmup 5 ) > gpurun_out/bench_warp.json 2> gpurun_out/bench_warp.err; cat gpurun_out/bench_warp.json | cut -c1-600
( timeout 200 python bench.py --steps 50 --warmup 5 --flow smooth --halo 8 --no-cpu-baseline ) > gpurun_out/bench_warp_smooth.json 2>> gpurun_out/bench_warp.err
( timeout 200 python bench.py --steps 50 --warmup 5 --method nearest --no-cpu-baseline ) > gpurun_out/bench_warp_nearest.json 2>> gpurun_out/bench_warp.err
for op in dice cce lc3d resize; do
  ( timeout 300 python bench.py --op $op --steps 20 --warmup 3 ) > gpurun_out/bench_$op.json 2> gpurun_out/bench_$op.err; cut -c1-300 gpurun_out/bench_$op.json
done
( timeout 300 python bench.py --op lc3d --lc-batch 8 --steps 5 --warmup 3 ) > gpurun_out/bench_lc3d_b8.json 2>> gpurun_out/bench_lc3d.err
( timeout 200 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/bench_reference.json 2>&1
( SWEEP_QUICK=0 timeout 600 python tools/sweep_warp.py ) > gpurun_out/sweep.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_warp.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:warp3d_tile -s 3 -c 1 -o gpurun_out/prof_warp -f \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_warp.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dice_sums -s 3 -c 1 -o gpurun_out/prof_dice -f \
  python bench.py --op dice --steps 2 --warmup 3 > gpurun_out/ncu_full_dice.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cce_vec4 -s 3 -c 1 -o gpurun_out/prof_cce -f \
  python bench.py --op cce --steps 2 --warmup 3 > gpurun_out/ncu_full_cce.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lc3d_stream -s 3 -c 1 -o gpurun_out/prof_lc3d -f \
  python bench.py --op lc3d --steps 2 --warmup 3 > gpurun_out/ncu_full_lc3d.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lc3d_stream -s 3 -c 1 -o gpurun_out/prof_lc3d_b8 -f \
  python bench.py --op lc3d --lc-batch 8 --steps 2 --warmup 3 > gpurun_out/ncu_full_lc3d_b8.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resize3d -s 3 -c 1 -o gpurun_out/prof_resize -f \
  python bench.py --op resize --steps 2 --warmup 3 > gpurun_out/ncu_full_resize.log 2>&1
ls -la gpurun_out | grep -E "prof|launches"
