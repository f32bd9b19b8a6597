#!/bin/bash
# round 2, multi-GPU visit: peer-memory transport of the slab plan (N = $1)
N=${1:-2}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_multi_gpu.py -q --timeout 600 -m gpu 2>&1 | tail -15 ) > gpurun_out/r2m${N}_pytest_multi.log 2>&1; tail -4 gpurun_out/r2m${N}_pytest_multi.log
run() {
  ( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N $2 ) > gpurun_out/r2m${N}_$1.json 2>> gpurun_out/r2m${N}.err
  python - "$1" "$N" <<'PY'
import json, sys
name, n = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open('gpurun_out/r2m%s_%s.json' % (n, name)).read().strip().splitlines()[-1])
    print(name, json.dumps({k: d[k] for k in d if k in ('overlap', 'overlap_nccl', 'serial', 'one_gpu_whole_volume', 'batch', 'slab', 'error')})[:2600])
except Exception as e:
    print(name, 'unreadable', e)
PY
}
run slab_c1 "--op warp_slab --slab-channels 1 --steps 200"
run slab_c16 "--op warp_slab --slab-channels 16 --steps 100"
run slab_c16_b8 "--op warp_slab --slab-channels 16 --slab-batch 8 --steps 50"
run cfg5 "--op cfg5 --cfg5-steps 5"
grep -v "^\*\|^$\|OMP_NUM" gpurun_out/r2m${N}.err | tail -12
