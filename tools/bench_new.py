#!/usr/bin/env python
"""Dev tool (GPU box): per-kernel timings of the MutualInformation and GaussianBlur paths."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neurite_b200 as ne  # noqa: E402
from neurite_b200 import utils  # noqa: E402
from neurite_b200._lib import lib, check, ptr, stream_ptr  # noqa: E402

S = (160, 192, 224)
V = S[0] * S[1] * S[2]
PEAK = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs'] if os.path.exists('MEASURED_PEAKS.json') else 6650.0


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


def report(name, ms, nbytes):
    print('%-58s %8.3f ms  %7.0f GB/s  frac %.3f' % (name, ms, nbytes / ms / 1e6, nbytes / ms / 1e6 / PEAK), flush=True)


dev = torch.device('cuda')
B = 8
x = torch.rand((B,) + S + (1,), device=dev)
y = (0.7 * x * x + 0.1 + 0.1 * torch.rand_like(x)).clamp_(0, 1)
if not os.environ.get('SKIP_MI'):
    m = ne.metrics.MutualInformation(nb_bins=16)
    report('mi.volumes whole call, B=8', timeit(lambda: m.volumes(x, y)), 8 * B * V)
    report('  minmax (one tensor)', timeit(lambda: utils.minmax(x)), 4 * B * V)
    cx = utils.bin_centers_from_range(utils.minmax(x), 16)
    cy = utils.bin_centers_from_range(utils.minmax(y), 16)
    xv, yv = x.reshape(B, V, 1), y.reshape(B, V, 1)
    for nb in (16, 32):
        mm = ne.metrics.MutualInformation(nb_bins=nb)
        ccx = utils.bin_centers_from_range(utils.minmax(x), nb)
        ccy = utils.bin_centers_from_range(utils.minmax(y), nb)
        stats = torch.empty((B, nb * nb + 2 * nb), device=dev)
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
        wsb = lib.nrt_mi_workspace_bytes(B, nb, nb)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)

        def hist():
            check(lib.nrt_mi_hist_f32(ptr(xv), V, 1, 1, nb, ptr(ccx), ptr(yv), V, 1, 1, nb, ptr(ccy), B, 1, V,
                                      float(mm.soft_bin_alpha), float('-inf'), float('inf'), ptr(stats), ptr(flag), ptr(ws), wsb,
                                      stream_ptr(dev)))
        report('  hist+combine quant/quant nb=%d (tensor cores)' % nb, timeit(hist), 8 * B * V)
        if nb == 16:
            for var in (1, 2):
                os.environ['NRT_MI_VARIANT'] = str(var)
                report('    variant %d' % var, timeit(hist), 8 * B * V)
            del os.environ['NRT_MI_VARIANT']
            for per_sm in (2, 8):
                os.environ['NRT_MI_CTAS_PER_SM'] = str(per_sm)
                report('    %d CTAs per SM' % per_sm, timeit(hist), 8 * B * V)
            del os.environ['NRT_MI_CTAS_PER_SM']
    xg, yg = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    out = m.volumes(xg, yg)
    go = torch.ones_like(out)

    def bwd():
        xg.grad = None
        yg.grad = None
        out.backward(go, retain_graph=True)
    report('mi.volumes backward (gx + gy), B=8', timeit(bwd, 5), 16 * B * V)
    del xg, yg, out
    del x, y, xv, yv
    for L, Bm in ((16, 2), (32, 1)):
        px = torch.softmax(torch.randn((Bm,) + S + (L,), device=dev), -1)
        py = torch.softmax(torch.randn((Bm,) + S + (L,), device=dev), -1)
        report('mi.segs whole call L=%d B=%d' % (L, Bm), timeit(lambda: m.segs(px, py)), 8 * L * Bm * V)
        del px, py
        torch.cuda.empty_cache()

x = torch.randn((B,) + S + (1,), device=dev)
for sigma in (1.0, 2.0, 4.0):
    lay = ne.layers.GaussianBlur(sigma=sigma)
    report('GaussianBlur sigma=%g whole call (3 passes), B=8, C=1' % sigma, timeit(lambda: lay(x)), 8 * B * V)
    k = utils.gaussian_kernel(sigma, device=dev)
    for ax in range(3):
        for mode in ((0, 1, 2) if ax < 2 else (1,)):
            os.environ['NRT_CONV_COL'] = str(mode)
            report('  pass axis %d (K=%d) col mode %d' % (ax, k.numel(), mode),
                   timeit(lambda: utils.separable_conv(x, k, axis=ax, batched=True)), 8 * B * V)
        del os.environ['NRT_CONV_COL']
x3 = torch.randn((4,) + S + (3,), device=dev)
lay = ne.layers.GaussianBlur(sigma=1.0)
report('GaussianBlur sigma=1 whole call, B=4, C=3', timeit(lambda: lay(x3)), 8 * 3 * 4 * V)
