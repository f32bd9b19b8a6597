#!/usr/bin/env python
"""Dev tool (GPU box), round 2: time the warp kernel variants on BASELINE.json's 160x192x224 volume.

  * C = 1: box-tile kernel, i.i.d. / smooth / zero flows, box following on / off
  * C = 2..32: generic gather vs box-tile (C <= 4) vs z-marching ring kernel (8 / 16 consumer warps),
    i.i.d. and smooth flows; every variant is compared bit for bit with the generic kernel first.
Prints ms per launch and the fraction of the measured HBM roofline ((12 + 8C) B per voxel)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neurite_b200 as ne  # noqa: E402,F401
from neurite_b200 import utils  # noqa: E402

S = (160, 192, 224)
V = S[0] * S[1] * S[2]
try:
    PEAK = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs']
except Exception:
    PEAK = 6650.0


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def setenv(env):
    for k in ('NRT_WARP_TILE', 'NRT_WARP_MARCH', 'NRT_MARCH_SMALLC', 'NRT_MARCH_NW', 'NRT_WARP_FOLLOW', 'NRT_MARCH_NSEG',
              'NRT_WARP_TILE_CFG', 'NRT_MARCH_QPT'):
        os.environ.pop(k, None)
    os.environ.update(env)


def flows_for(B, dev, g):
    fl = {'iid3': torch.rand((B,) + S + (3,), device=dev, generator=g) * 6 - 3}
    coarse = torch.randn((B, 3, 10, 12, 14), device=dev, generator=g)
    sm = torch.nn.functional.interpolate(coarse, size=S, mode='trilinear', align_corners=True)
    fl['smooth3'] = (sm / sm.abs().amax() * 3).permute(0, 2, 3, 4, 1).contiguous()
    return fl


def main():
    dev = torch.device('cuda')
    g = torch.Generator(device=dev).manual_seed(0)
    only = os.environ.get('SWEEP_ONLY', '')
    if only in ('', 'c1', 'quick'):
        B = 8
        vol = torch.randn((B,) + S + (1,), device=dev, generator=g)
        fl = flows_for(B, dev, g)
        fl['zero'] = torch.zeros_like(fl['iid3'])
        coarse = torch.randn((B, 3, 10, 12, 14), device=dev, generator=g)
        sm = torch.nn.functional.interpolate(coarse, size=S, mode='trilinear', align_corners=True)
        fl['smooth8'] = (sm / sm.abs().amax() * 8).permute(0, 2, 3, 4, 1).contiguous()
        for fname, flow in fl.items():
            for method in ('linear', 'nearest'):
                for label, env in (('tile follow', {}), ('tile nofollow', {'NRT_WARP_FOLLOW': '0'})):
                    setenv(env)
                    ms = timeit(lambda: utils._warp_batched(vol, flow, method, None), 50)
                    gbs = 20.0 * B * V / ms / 1e6
                    print('C=1  B=%d %-8s %-7s %-14s: %.4f ms  %.3e vox/s  frac %.3f' %
                          (B, fname, method, label, ms, B * V / ms * 1e3, gbs / PEAK), flush=True)
        del vol, fl
    for C in (16, 8, 4, 3, 2, 32) if only != 'quick' else (16, 8):
        if only not in ('', 'multi', 'quick', 'c%d' % C):
            continue
        B = max(1, min(8, 32 // C))
        vol = torch.randn((B,) + S + (C,), device=dev, generator=g)
        for fname, flow in flows_for(B, dev, g).items():
            variants = [('generic', {'NRT_WARP_TILE': '0', 'NRT_WARP_MARCH': '0'})]
            if C <= 4:
                variants.append(('box-tile', {'NRT_WARP_MARCH': '0'}))
            variants += [('march nw16', {'NRT_MARCH_SMALLC': '1'}), ('march nw8', {'NRT_MARCH_SMALLC': '1', 'NRT_MARCH_NW': '8'})]
            if C % 8 == 0:
                variants.append(('march qpt2 nw8', {'NRT_MARCH_QPT': '2'}))
            if C % 16 == 0:
                variants.append(('march qpt4 nw4', {'NRT_MARCH_QPT': '4'}))
            ref = None
            for method in ('linear',) if fname != 'iid3' else ('linear', 'nearest'):
                for label, env in variants:
                    setenv(env)
                    out = utils._warp_batched(vol[:1], flow[:1], method, None)
                    if label == 'generic':
                        ref = out
                    same = bool(torch.equal(out, ref))
                    ms = timeit(lambda: utils._warp_batched(vol, flow, method, None), 10)
                    gbs = (12.0 + 8.0 * C) * B * V / ms / 1e6
                    print('C=%-2d B=%d %-8s %-7s %-14s: %.4f ms  %.3e vox/s  frac %.3f  equal_to_generic=%s' %
                          (C, B, fname, method, label, ms, B * V / ms * 1e3, gbs / PEAK, same), flush=True)
        del vol
    setenv({})


if __name__ == '__main__':
    main()
