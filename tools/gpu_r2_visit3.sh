#!/bin/bash
# round 2, visit 3 (1 GPU): LC3D aliasing fix -- the configurations that timed out before, then the whole suite
mkdir -p gpurun_out
cat > /tmp/lc_b2.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from neurite_b200.layers import local_conv3d
I = int(sys.argv[1]); B = int(sys.argv[2])
O = I - 2
x = torch.randn((B, I, I, I, 16), device='cuda')
k = torch.rand((O ** 3, 432, 16), device='cuda') * 0.1
try:
    for _ in range(5):
        y = local_conv3d(x, k, None, (3, 3, 3), (1, 1, 1), (O, O, O))
    torch.cuda.synchronize()
    import os
    os.environ['NRT_LC3D_PATCH'] = '0'
    ref = local_conv3d(x, k, None, (3, 3, 3), (1, 1, 1), (O, O, O))
    print('ok equal_to_stream_kernel=%s' % bool(torch.equal(y, ref)))
except Exception as e:
    print('FAILED', str(e)[:80])
PY
for cfg in 12 21; do for nw in 3 6 7; do echo "== B2 cfg $cfg warps $nw"; NRT_LC3D_B2=$cfg NRT_LC3D_WARPS=$nw timeout 60 python /tmp/lc_b2.py 64 2 2>&1 | tail -1; done; done
for st in 4 5; do echo "== stages $st"; NRT_LC3D_STAGES=$st timeout 60 python /tmp/lc_b2.py 64 2 2>&1 | tail -1; done
for b in 3 5 8; do echo "== B=$b"; timeout 60 python /tmp/lc_b2.py 64 $b 2>&1 | tail -1; done
( timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -12 ) > gpurun_out/r2v3_pytest_all.log 2>&1
tail -6 gpurun_out/r2v3_pytest_all.log
for b in 2 4 8; do ( timeout 300 python bench.py --op lc3d --lc-batch $b --no-cpu-baseline ) > gpurun_out/r2v3_lc3d_b$b.json 2>> gpurun_out/r2v3.err; python -c "
import json; d=json.loads(open('gpurun_out/r2v3_lc3d_b$b.json').read().strip().splitlines()[-1]); print('lc3d B=$b', d['ms_per_step'], d['roofline']['frac'])"; done
( timeout 300 python bench.py --op lc3d --no-cpu-baseline ) > gpurun_out/r2v3_lc3d_b1.json 2>> gpurun_out/r2v3.err; python -c "
import json; d=json.loads(open('gpurun_out/r2v3_lc3d_b1.json').read().strip().splitlines()[-1]); print('lc3d B=1', d['ms_per_step'], d['roofline']['frac'])"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lc3d_patch -s 2 -c 1 -o gpurun_out/r2v3_prof_lc3d_b8 -f python bench.py --op lc3d --lc-batch 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2v3_ncu_lc3d.log 2>&1
tail -3 gpurun_out/r2v3.err
