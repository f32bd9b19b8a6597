#!/bin/bash
# round 2, last 1-GPU visit: the round-end sequence the driver runs (GPU tests, smoke, reference arm, default bench), a
# fresh launch list of the default bench command and one capture each of the kernels that changed last (CCE, Dice combine)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6 ) > gpurun_out/r2z_pytest_all.log 2>&1; tail -3 gpurun_out/r2z_pytest_all.log
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/r2z_smoke.log 2>&1; tail -1 gpurun_out/r2z_smoke.log
( timeout 300 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r2z_bench_reference.json 2>> gpurun_out/r2z_bench.err; cut -c1-300 gpurun_out/r2z_bench_reference.json
( timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2z_bench_default.json 2>> gpurun_out/r2z_bench.err; cut -c1-400 gpurun_out/r2z_bench_default.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z_bench_default.json').read().strip().splitlines()[-1])
for k,v in d.get('ops',{}).items(): print(k, v.get('ms_per_step'), (v.get('roofline') or {}).get('frac'), (v.get('clocks') or {}).get('sm_mhz'), v.get('error'))
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2z_launches_bench.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2z_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cce_vec4u -s 3 -c 1 -o gpurun_out/r2z_prof_cce -f \
    python bench.py --op cce --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2z_ncu_full_cce.log 2>&1
ls -la gpurun_out | grep r2z_ | wc -l
