#!/bin/bash
# multi-GPU visit (gpurun --gpus N): sharded == single tests, scaling bench lines
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
( timeout 600 python -m pytest tests/test_multi_gpu.py -q --timeout 400 -m gpu 2>&1 | tail -15 ) > gpurun_out/pytest_multi_gpu.log 2>&1; tail -3 gpurun_out/pytest_multi_gpu.log
for n in 1 $N; do
  if [ "$n" = "1" ]; then
    ( timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline ) > gpurun_out/scale_warp_n1.json 2> gpurun_out/scale.err
  else
    ( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 30 --warmup 5 --no-cpu-baseline ) > gpurun_out/scale_warp_n$n.json 2>> gpurun_out/scale.err
    ( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --op dice --steps 20 --warmup 5 ) > gpurun_out/scale_dice_n$n.json 2>> gpurun_out/scale.err
  fi
  python -c "import json,sys; d=json.loads(open('gpurun_out/scale_warp_n$n.json').read().strip().splitlines()[-1]); print('warp n=$n', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
done
python -c "import json; d=json.loads(open('gpurun_out/scale_dice_n$N.json').read().strip().splitlines()[-1]); print('dice n=$N', d['value'], d['ms_per_step'], d['roofline']['frac'])"
( timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 3 --warmup 1 ) > gpurun_out/scale_ref_n$N.json 2>> gpurun_out/scale.err; cut -c1-200 gpurun_out/scale_ref_n$N.json
tail -5 gpurun_out/scale.err
