#!/usr/bin/env python
"""Dev tool (GPU box): time the warp kernel variants on BASELINE.json's 160x192x224 volume.
Prints ms per batch-8 launch, voxels/s and fraction of the measured HBM roofline."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neurite_b200 as ne  # noqa: E402
from neurite_b200 import utils  # noqa: E402

S = (160, 192, 224)
V = S[0] * S[1] * S[2]
B = int(os.environ.get('SWEEP_BATCH', '8'))
try:
    PEAK = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs']
except Exception:
    PEAK = 6650.0


def timeit(fn, iters=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device('cuda')
    g = torch.Generator(device=dev).manual_seed(0)
    vol = torch.randn((B,) + S + (1,), device=dev, generator=g)
    flows = {'iid3': torch.rand((B,) + S + (3,), device=dev, generator=g) * 6 - 3}
    coarse = torch.randn((B, 3, 10, 12, 14), device=dev, generator=g)
    sm = torch.nn.functional.interpolate(coarse, size=S, mode='trilinear', align_corners=True)
    flows['smooth8'] = (sm / sm.abs().amax() * 8).permute(0, 2, 3, 4, 1).contiguous()
    flows['zero'] = torch.zeros_like(flows['iid3'])
    # reference point: a plain device copy of the same 20 B/voxel
    a = torch.empty(B * V * 5 // 2, device=dev)
    b = torch.empty_like(a)
    ms = timeit(lambda: b.copy_(a))
    print('copy 20B/vox          : %.3f ms  %.0f GB/s' % (ms, 20.0 * B * V / ms / 1e6))
    quick = os.environ.get('SWEEP_QUICK') == '1'
    checked = set()
    for fname, flow in flows.items():
        if quick and fname != 'iid3':
            continue
        for method in ('linear', 'nearest'):
            for label, env in [('generic', {'NRT_WARP_TILE': '0'})] + \
                              [('tile cfg%d halo%d' % (c, h), {'NRT_WARP_TILE': '1', 'NRT_WARP_TILE_CFG': str(c), 'H': h})
                               for c in (0, 2, 3, 4, 5) for h in ((3, 4) if fname != 'smooth8' else (3, 4, 6, 8))
                               if not (c in (4, 5) and h > 4)]:
                h = env.pop('H', 0)
                if quick and not (label.startswith('tile cfg') and label.endswith('halo3') and label[8] in '02345'):
                    continue
                os.environ['NRT_WARP_PERSIST'] = env.get('NRT_WARP_PERSIST', '0')
                os.environ.update(env)
                if label[5:9] in ('cfg4', 'cfg5') and (fname, method) not in checked:
                    # the occupancy-capped variants are not in the test-suite: check them against the default here
                    os.environ['NRT_WARP_TILE_CFG'] = '2'
                    ref = utils._warp_batched(vol[:1], flow[:1], method, None, halo=h)
                    os.environ.update(env)
                    ok = torch.equal(ref, utils._warp_batched(vol[:1], flow[:1], method, None, halo=h))
                    print('%-8s %-7s %-18s: equal to cfg2: %s' % (fname, method, label, ok))
                    if label[5:9] == 'cfg5':
                        checked.add((fname, method))
                ms = timeit(lambda: utils._warp_batched(vol, flow, method, None, halo=h))
                gbs = 20.0 * B * V / ms / 1e6
                print('%-8s %-7s %-18s: %.3f ms  %.3e vox/s  %.0f GB/s  frac %.3f' %
                      (fname, method, label, ms, B * V / ms * 1e3, gbs, gbs / PEAK))
            sys.stdout.flush()


if __name__ == '__main__':
    main()
