#!/bin/bash
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
    for _ in range(2):
        y = local_conv3d(x, k, None, (3, 3, 3), (1, 1, 1), (O, O, O))
    torch.cuda.synchronize()
    print('ok', float(y.abs().sum()))
except Exception as e:
    print('FAILED', str(e)[:80])
PY
echo "== cfg 12 warps 3 I=64 (fails): timeout report"; NRT_LC3D_B2=12 NRT_LC3D_WARPS=3 timeout 60 python /tmp/lc_b2.py 64 2 2>&1 | grep -v "^$" | head -12
echo "== cfg 12 warps 3 stages 5"; NRT_LC3D_STAGES=5 NRT_LC3D_B2=12 NRT_LC3D_WARPS=3 timeout 60 python /tmp/lc_b2.py 64 2 2>&1 | grep -v "^$" | head -6
echo "== cfg 12 warps 3 stages 4"; NRT_LC3D_STAGES=4 NRT_LC3D_B2=12 NRT_LC3D_WARPS=3 timeout 60 python /tmp/lc_b2.py 64 2 2>&1 | grep -v "^$" | head -6
echo "== cfg 12 default warps, stages 5"; NRT_LC3D_STAGES=5 NRT_LC3D_B2=12 timeout 60 python /tmp/lc_b2.py 64 2 2>&1 | grep -v "^$" | head -6
echo "== racecheck cfg 12 I=16"; NRT_LC3D_B2=12 timeout 250 compute-sanitizer --tool racecheck --print-limit 4 python /tmp/lc_b2.py 16 2 2>&1 | grep -v "^$" | head -40
