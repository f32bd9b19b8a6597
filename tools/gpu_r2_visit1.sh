#!/bin/bash
# round 2, visit 1 (1 GPU): new march kernel + reworked box-tile kernel: correctness first, then the sweep, then ncu
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "march or tile_configs or follows or batch_and_channels or slabs or full_size" 2>&1 | tail -25 ) > gpurun_out/r2v1_pytest_warp.log 2>&1
tail -4 gpurun_out/r2v1_pytest_warp.log
( timeout 900 python tools/sweep_r2.py ) > gpurun_out/r2v1_sweep.txt 2>&1; tail -90 gpurun_out/r2v1_sweep.txt
( timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -25 ) > gpurun_out/r2v1_pytest_all.log 2>&1
tail -4 gpurun_out/r2v1_pytest_all.log
( timeout 300 python bench.py --steps 200 --warmup 5 --no-numpy-baseline ) > gpurun_out/r2v1_bench_warp.json 2> gpurun_out/r2v1_bench_warp.err; cut -c1-600 gpurun_out/r2v1_bench_warp.json
# ncu: march kernel (C=16) and box-tile kernel (C=1), one launch each
cat > /tmp/one_warp.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from neurite_b200 import utils
C, B = int(sys.argv[1]), int(sys.argv[2])
S = (160, 192, 224)
vol = torch.randn((B,) + S + (C,), device='cuda')
flow = torch.rand((B,) + S + (3,), device='cuda') * 6 - 3
for _ in range(3):
    utils._warp_batched(vol, flow)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:warp3d_march -s 2 -c 1 -o gpurun_out/r2v1_prof_march16 -f python /tmp/one_warp.py 16 2 > gpurun_out/r2v1_ncu_march.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:warp3d_tile -s 2 -c 1 -o gpurun_out/r2v1_prof_tile1 -f python /tmp/one_warp.py 1 8 > gpurun_out/r2v1_ncu_tile.log 2>&1
ls -la gpurun_out | grep r2v1
