#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
( timeout 900 python -m pytest tests/test_multi_gpu.py -q --timeout 600 -m gpu 2>&1 | tail -8 ) > gpurun_out/pytest_multi_gpu8.log 2>&1; tail -3 gpurun_out/pytest_multi_gpu8.log
for n in 1 2 4 8; do
  if [ "$n" = "1" ]; then
    ( timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline ) > gpurun_out/scale_warp_n1.json 2> gpurun_out/scale.err
  else
    ( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 30 --warmup 5 --no-cpu-baseline ) > gpurun_out/scale_warp_n$n.json 2>> gpurun_out/scale.err
  fi
  python -c "import json,sys; d=json.loads(open('gpurun_out/scale_warp_n$n.json').read().strip().splitlines()[-1]); print('warp n=$n', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
done
for n in 2 8; do
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --op dice --steps 20 --warmup 5 ) > gpurun_out/scale_dice_n$n.json 2>> gpurun_out/scale.err
python -c "import json; d=json.loads(open('gpurun_out/scale_dice_n$n.json').read().strip().splitlines()[-1]); print('dice n=$n', d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
grep -i -E "error|Traceback" gpurun_out/scale.err | head -5
