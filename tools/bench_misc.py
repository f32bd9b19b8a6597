#!/usr/bin/env python
"""Dev tool (GPU box): timings of the paths bench.py does not cover -- multi-channel warp,
backward kernels, hard Dice -- with their algorithmic-byte rooflines."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neurite_b200 as ne  # noqa: E402

S = (160, 192, 224)
V = S[0] * S[1] * S[2]
PEAK = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs'] if os.path.exists('MEASURED_PEAKS.json') else 6650.0


def timeit(fn, iters=10):
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


def report(name, ms, nbytes):
    print('%-46s %8.3f ms  %7.0f GB/s  frac %.3f' % (name, ms, nbytes / ms / 1e6, nbytes / ms / 1e6 / PEAK), flush=True)


dev = torch.device('cuda')
g = torch.Generator(device=dev).manual_seed(0)
for C, B in ((1, 8), (2, 4), (3, 4), (4, 4), (16, 2)):
    vol = torch.randn((B,) + S + (C,), device=dev, generator=g)
    flow = torch.rand((B,) + S + (3,), device=dev, generator=g) * 6 - 3
    st = ne.layers.SpatialTransformer()
    report('warp fwd C=%d B=%d' % (C, B), timeit(lambda: st([vol, flow])), (12 + 8 * C) * B * V)
    v = vol.clone().requires_grad_(True)
    f = flow.clone().requires_grad_(True)
    out = st([v, f])
    go = torch.randn_like(out)

    def bwd():
        v.grad = None
        f.grad = None
        out.backward(go, retain_graph=True)
    # bytes: read vol-corners (C*4 once), flow 12, gout 4C; write gvol 4C (+zero fill 4C), gflow 12
    report('warp bwd C=%d B=%d (gvol+gflow)' % (C, B), timeit(bwd), (24 + 16 * C) * B * V)
    del vol, flow, v, f, out, go
    torch.cuda.empty_cache()

B, L = 4, 16
lab = torch.randint(0, L, (B,) + S, device=dev, generator=g)
t = torch.nn.functional.one_hot(lab, L).float()
p = torch.softmax(torch.randn((B,) + S + (L,), device=dev, generator=g), -1).requires_grad_(True)
loss = ne.losses.Dice().mean_loss(t, p)


def dbwd():
    p.grad = None
    loss.backward(retain_graph=True)


report('dice bwd [4,160,192,224,16]', timeit(dbwd), 12.0 * B * V * L)
closs = ne.losses.CategoricalCrossentropy().loss(t, p)


def cbwd():
    p.grad = None
    closs.backward(retain_graph=True)


report('cce bwd [4,160,192,224,16]', timeit(cbwd), 12.0 * B * V * L)
hd = ne.losses.HardDice(L)
lp = torch.randint(0, L, (B,) + S, device=dev, generator=g).int()
lt = lab.int()
report('hard dice (labels) [4,160,192,224]', timeit(lambda: hd.loss(lt, lp)), 8.0 * B * V)
