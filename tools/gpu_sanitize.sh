#!/bin/bash
# compute-sanitizer pass over the small-shape GPU tests (memcheck + racecheck on the kernels
# that use shared memory / mbarriers).  Summary goes to gpurun_out/sanitizer.txt
mkdir -p gpurun_out
{
echo "== memcheck: tile warp (all configs, 1-4 channels), resize, slabs, LC3D, Dice/CCE, gradients"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_grad.py -q -x --timeout 600 \
   -k "tile_configs or batch_and_channels or follows or slabs or lc3d_golden or dice_golden or cce_golden or gradient or empty or vxm" 2>&1 | tail -6
echo "== racecheck: tile warp + backward tile + LC3D ring"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_grad.py -q -x --timeout 600 \
   -k "follows or lc3d_golden or (gradient and shape4)" 2>&1 | tail -6
} > gpurun_out/sanitizer.txt 2>&1
cat gpurun_out/sanitizer.txt
