"""
GPU gradient tests.  The reference gets its gradients from TensorFlow
autodiff of the op graph; here the same graph is written in differentiable torch float64
(floor: zero grad; clamp: passes on the closed interval, like tf.clip_by_value; gather:
scatter-add) and torch.autograd provides the expected gradients.
"""
import itertools

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ne(cuda):
    import neurite_b200
    return neurite_b200


def interpn_torch(vol, loc, method='linear', fill=None):
    """reference utils.py:137-213 in differentiable torch (vol [*S,C], loc [*O,D])."""
    D = loc.shape[-1]
    S = vol.shape[:-1]
    mx = [s - 1 for s in S]
    flat = vol.reshape(-1, vol.shape[-1])
    strides = [int(np.prod(S[d + 1:])) for d in range(D)]
    if method == 'linear':
        loc0 = torch.floor(loc)
        x = [torch.clamp(loc[..., d], 0, mx[d]) for d in range(D)]
        f0 = [torch.clamp(loc0[..., d], 0, mx[d]) for d in range(D)]
        f1 = [torch.clamp(f0[d] + 1, 0, mx[d]) for d in range(D)]
        idx = [[f.long() for f in f0], [f.long() for f in f1]]
        wlo = [f1[d].detach() - x[d] for d in range(D)]
        w = [wlo, [1 - v for v in wlo]]
        out = 0
        for c in itertools.product([0, 1], repeat=D):
            ii = sum(idx[c[d]][d] * strides[d] for d in range(D))
            ww = w[c[0]][0]
            for d in range(1, D):
                ww = ww * w[c[d]][d]
            out = out + ww[..., None] * flat[ii]
    else:
        r = [torch.clamp(torch.round(loc[..., d]).long(), 0, mx[d]) for d in range(D)]
        out = flat[sum(r[d] * strides[d] for d in range(D))]
    if fill is not None:
        oob = torch.zeros_like(loc[..., 0], dtype=torch.bool)
        for d in range(D):
            oob = oob | (loc[..., d] < 0) | (loc[..., d] > mx[d])
        out = out * (~oob)[..., None].to(out.dtype) + oob[..., None].to(out.dtype) * fill
    return out


@pytest.mark.parametrize('shape,C', [((6, 7, 8), 1), ((5, 6, 9), 3), ((9, 10), 2), ((17,), 1), ((10, 12, 36), 1),
                                     ((19, 9, 64), 1)])
@pytest.mark.parametrize('method,fill', [('linear', None), ('linear', 0.5), ('nearest', None)])
def test_warp_and_interpn_gradients(ne, shape, C, method, fill):
    g = torch.Generator().manual_seed(len(shape) * 10 + C)
    D = len(shape)
    vol = torch.randn((2,) + shape + (C,), generator=g, dtype=torch.float64)
    # keep sample points away from integer coordinates (kinks) except a few exact edges
    flow = (torch.rand((2,) + shape + (D,), generator=g, dtype=torch.float64) * 5 - 2.5)
    grid = torch.stack(torch.meshgrid(*[torch.arange(s, dtype=torch.float64) for s in shape], indexing='ij'), -1)
    loc = grid[None] + flow
    frac = loc - torch.floor(loc)
    flow = flow + ((frac < 0.05).double() * 0.1 - (frac > 0.95).double() * 0.1)
    flow = flow.float().double()
    gout = torch.randn((2,) + shape + (C,), generator=g, dtype=torch.float64).float().double()
    vol = vol.float().double()

    v_ref = vol.clone().requires_grad_(True)
    f_ref = flow.clone().requires_grad_(True)
    out_ref = torch.stack([interpn_torch(v_ref[b], grid + f_ref[b], method, fill) for b in range(2)])
    out_ref.backward(gout)

    v = vol.float().cuda().requires_grad_(True)
    f = flow.float().cuda().requires_grad_(True)
    out = ne.layers.SpatialTransformer(interp_method=method, fill_value=fill)([v, f])
    np.testing.assert_allclose(out.detach().cpu().numpy(), out_ref.detach().numpy(), rtol=1e-5, atol=1e-5)
    out.backward(gout.float().cuda())
    np.testing.assert_allclose(v.grad.cpu().numpy(), v_ref.grad.numpy(), rtol=1e-4, atol=1e-4)
    f_exp = f_ref.grad.numpy() if f_ref.grad is not None else np.zeros(f_ref.shape)   # nearest: no gradient to loc
    np.testing.assert_allclose(f.grad.cpu().numpy(), f_exp, rtol=1e-4, atol=1e-4)

    # interpn with an explicit loc tensor
    v2 = vol[0].float().cuda().requires_grad_(True)
    l2 = (grid + flow[0]).float().cuda().requires_grad_(True)
    o2 = ne.utils.interpn(v2, l2, method, fill)
    o2.backward(gout[0].float().cuda())
    v3 = vol[0].clone().requires_grad_(True)
    l3 = (grid + flow[0]).float().double().requires_grad_(True)
    interpn_torch(v3, l3, method, fill).backward(gout[0])
    np.testing.assert_allclose(v2.grad.cpu().numpy(), v3.grad.numpy(), rtol=1e-4, atol=1e-4)
    l_exp = l3.grad.numpy() if l3.grad is not None else np.zeros(l3.shape)
    np.testing.assert_allclose(l2.grad.cpu().numpy(), l_exp, rtol=1e-4, atol=1e-4)


def test_resize_gradient(ne):
    g = torch.Generator().manual_seed(3)
    x = torch.randn((2, 5, 6, 7, 3), generator=g)
    xg = x.cuda().requires_grad_(True)
    out = ne.layers.Resize([2, 1.5, 0.8])(xg)
    gout = torch.randn(out.shape, generator=g)
    out.backward(gout.cuda())
    # expected: the same sampling expressed through interpn_torch on the linspace grid
    xr = x.double().requires_grad_(True)
    M = out.shape[1:-1]
    lin = []
    for d in range(3):
        S, m = x.shape[1 + d], M[d]
        delta = np.float32(S - 1) / np.float32(m - 1)
        v = (delta * np.arange(m, dtype=np.float32)).astype(np.float32)
        v[-1] = S - 1
        lin.append(torch.from_numpy(v).double())
    grid = torch.stack(torch.meshgrid(*lin, indexing='ij'), -1)
    ref = torch.stack([interpn_torch(xr[b], grid) for b in range(2)])
    ref.backward(gout.double())
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref.detach().numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(xg.grad.cpu().numpy(), xr.grad.numpy(), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize('laplace', [0.0, 0.1])
def test_dice_gradient(ne, laplace):
    g = torch.Generator().manual_seed(5)
    L = 5
    lab = torch.randint(0, 4, (2, 6, 7, 8), generator=g)              # label 4 absent: 0/0 -> 0, zero gradient
    t = torch.nn.functional.one_hot(lab, L).double()
    p = torch.softmax(torch.randn((2, 6, 7, 8, L), generator=g, dtype=torch.float64), -1)
    p[..., 4] = 0
    w = torch.rand((1, L), generator=g, dtype=torch.float64)
    pr = p.clone().requires_grad_(True)
    top = 2 * (t * pr).flatten(1, 3).sum(1)
    bot = (t * t).flatten(1, 3).sum(1) + (pr * pr).flatten(1, 3).sum(1)
    dice = (top + laplace) / (bot + laplace) if laplace > 0 else torch.where(bot != 0, top / torch.where(bot != 0, bot, torch.ones_like(bot)), torch.zeros_like(top))
    (-(dice * w).mean()).backward()
    pg = p.float().cuda().requires_grad_(True)
    loss = ne.losses.Dice(weights=w.float().numpy(), laplace_smoothing=laplace).mean_loss(t.float().cuda(), pg)
    loss.backward()
    np.testing.assert_allclose(float(loss), float(-(dice * w).mean()), rtol=1e-5)
    np.testing.assert_allclose(pg.grad.cpu().numpy(), pr.grad.numpy(), rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize('kw', [dict(), dict(from_logits=True), dict(label_smoothing=0.1), dict(reduction='sum'),
                                dict(reduction='none')])
def test_cce_gradient(ne, kw):
    g = torch.Generator().manual_seed(7)
    C = 6
    t = torch.nn.functional.one_hot(torch.randint(0, C, (2, 5, 6), generator=g), C).double()
    p = torch.rand((2, 5, 6, C), generator=g, dtype=torch.float64) + 0.05
    p[0, 0, 0] = torch.tensor([1.0, 0, 0, 0, 0, 0])                  # clipped entries: gradient masked
    lw = torch.rand(C, generator=g, dtype=torch.float64) + 0.5
    sw = torch.rand((2, 5, 6), generator=g, dtype=torch.float64) + 0.5
    pr = p.clone().requires_grad_(True)
    tt = t * lw
    ls = kw.get('label_smoothing', 0.0)
    if ls:
        tt = tt * (1 - ls) + ls / C
    if kw.get('from_logits'):
        l = -(tt * torch.log_softmax(pr, -1)).sum(-1)
    else:
        q = pr / pr.sum(-1, keepdim=True)
        l = -(tt * torch.log(torch.clamp(q, 1e-7, 1 - 1e-7))).sum(-1)
    l = l * sw
    red = kw.get('reduction', 'sum_over_batch_size')
    gper = torch.rand((2, 5, 6), generator=g, dtype=torch.float64)
    if red == 'none':
        l.backward(gper)
    else:
        (l.sum() if red == 'sum' else l.mean()).backward()
    pg = p.float().cuda().requires_grad_(True)
    c = ne.losses.CategoricalCrossentropy(label_weights=lw.float().numpy(), **kw)
    out = c(t.float().cuda(), pg, sample_weight=sw.float().cuda())
    if red == 'none':
        out.backward(gper.float().cuda())
    else:
        out.backward()
    np.testing.assert_allclose(pg.grad.cpu().numpy(), pr.grad.numpy(), rtol=2e-4, atol=1e-6)


def test_lc3d_gradient(ne):
    """LocallyConnected3D gradients vs torch autograd of an unfold + einsum formulation."""
    g = torch.Generator().manual_seed(9)
    B, I, Cin, Cout = 3, (6, 7, 8), 4, 8
    x = torch.randn((B,) + I + (Cin,), generator=g)
    lay = ne.layers.LocallyConnected3D(Cout, 3, activation='tanh')
    lay.build(x.shape)
    O = (lay.output_row, lay.output_col, lay.output_z)
    with torch.no_grad():
        lay.kernel.normal_(0, 0.2, generator=g)
        lay.bias.normal_(0, 1, generator=g)
    k_ref = lay.kernel.detach().double().clone().requires_grad_(True)
    b_ref = lay.bias.detach().double().clone().requires_grad_(True)
    x_ref = x.double().clone().requires_grad_(True)
    win = x_ref.unfold(1, 3, 1).unfold(2, 3, 1).unfold(3, 3, 1)            # [B,o0,o1,o2,Cin,k0,k1,k2]
    patches = win.permute(0, 1, 2, 3, 5, 6, 7, 4).reshape(B, -1, 27 * Cin)  # j = ((i0*3+i1)*3+i2)*Cin + c
    ref = torch.tanh(torch.einsum('bpj,pjf->bpf', patches, k_ref).reshape((B,) + O + (Cout,)) + b_ref)
    gout = torch.randn(ref.shape, generator=g, dtype=torch.float64)
    ref.backward(gout)
    lay.cuda()
    xg = x.cuda().requires_grad_(True)
    out = lay(xg)
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref.detach().numpy(), rtol=1e-4, atol=1e-5)
    out.backward(gout.float().cuda())
    np.testing.assert_allclose(xg.grad.cpu().numpy(), x_ref.grad.numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(lay.kernel.grad.cpu().numpy(), k_ref.grad.numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(lay.bias.grad.cpu().numpy(), b_ref.grad.numpy(), rtol=1e-4, atol=1e-4)
