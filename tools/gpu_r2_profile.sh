#!/bin/bash
# round 2, profile visit (1 GPU): the launch list of the default bench command + one `--set full` capture per hot kernel
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6 ) > gpurun_out/r2p_pytest_all.log 2>&1; tail -3 gpurun_out/r2p_pytest_all.log
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/r2p_smoke.log 2>&1; tail -1 gpurun_out/r2p_smoke.log
( timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2p_bench_default.json 2> gpurun_out/r2p_bench.err; cut -c1-400 gpurun_out/r2p_bench_default.json
( timeout 300 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r2p_bench_reference.json 2>> gpurun_out/r2p_bench.err; cut -c1-300 gpurun_out/r2p_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2p_launches_bench.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2p_ncu_launch.log 2>&1
prof() {  # name, kernel regex, bench args
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -o gpurun_out/r2p_prof_$1 -f \
    python bench.py $3 --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2p_ncu_full_$1.log 2>&1
}
prof warp warp3d_tile "--e2e-steps 1"
prof warp_c16 warp3d_march "--op warp_mc --channels 16"
prof resize resize3d "--op resize"
prof lc3d lc3d_patch "--op lc3d"
prof lc3d_b8 lc3d_rows "--op lc3d --lc-batch 8"
prof dice dice_sums "--op dice"
prof cce cce_vec4 "--op cce"
for op in "warp_mc --channels 16 --flow smooth" "warp_mc --channels 3" "warp_mc --channels 4" "lc3d --lc-batch 8" "lc3d --lc-batch 2" "mi" "blur" "resize"; do
  n=$(echo $op | tr ' -' '__'); ( timeout 300 python bench.py --op $op --no-cpu-baseline ) > gpurun_out/r2p_op_$n.json 2>> gpurun_out/r2p_bench.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2p_op_$n.json').read().strip().splitlines()[-1]); print('$op', d['ms_per_step'], d['roofline']['frac'])"
done
( timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 500 \
    -k "march_kernel_multichannel_bit_exact and 16-shape0 or resize_upsampling_vs_oracle and shape2 or lc3d_golden or warp_tile_configs_bit_exact and 2-shape0 or warp_slabs" 2>&1 | tail -8 ) > gpurun_out/r2p_sanitizer.txt 2>&1
tail -3 gpurun_out/r2p_sanitizer.txt
ls -la gpurun_out | grep r2p_ | wc -l
