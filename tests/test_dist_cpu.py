"""
CPU (gloo, world_size 2 and 3) tests of the multi-GPU host logic in neurite_b200.dist: slab
bounds, halo windows, the neighbour halo exchange and the source all-gather.  The compute
kernels are not called here (they need a GPU); the exchanged windows are checked against
plain slicing of the full volume, which is exactly what the kernel consumes.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neurite_b200 import dist as nd


def test_slab_bounds_cover_range_exactly():
    for n in (1, 7, 20, 160, 161):
        for w in (1, 2, 3, 8):
            b = nd.all_slab_bounds(n, w)
            assert b[0][0] == 0 and sum(c for _, c in b) == n
            for r in range(1, w):
                assert b[r][0] == b[r - 1][0] + b[r - 1][1]
            assert max(c for _, c in b) - min(c for _, c in b) <= 1
    assert nd.all_slab_bounds(160, 8) == [(20 * r, 20) for r in range(8)]       # cfg 5: 20 planes per GPU
    assert nd.source_window(20, 20, 4, 160) == (16, 44)
    assert nd.source_window(0, 20, 4, 160) == (0, 24)
    assert nd.source_window(140, 20, 4, 160) == (136, 160)
    f = torch.zeros(1, 4, 5, 6, 3)
    f[0, 1, 2, 3, 0] = -2.25
    f[0, 0, 0, 0, 1] = 9.0                     # in-plane shifts do not widen the z window
    assert nd.required_halo(f) == 4


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, full_s0, halo, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        full = torch.randn(2, full_s0, 5, 6, 1, generator=g)
        z0, nz = nd.slab_bounds(full_s0, world, rank)
        slab = full[:, z0:z0 + nz].contiguous()
        ext, src_z0 = nd.exchange_halo(slab, halo, full_s0)
        lo, hi = nd.source_window(z0, nz, halo, full_s0)
        ok = (src_z0 == lo) and torch.equal(ext, full[:, lo:hi])
        gathered = nd.gather_source(slab, full_s0)
        ok = ok and torch.equal(gathered, full)
        # Dice-style partial-sum all-reduce over voxel-range shards == unsharded sum
        part = slab.double().pow(2).sum().reshape(1)
        dist.all_reduce(part)
        ok = ok and bool(torch.allclose(part, full.double().pow(2).sum().reshape(1)))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,full_s0,halo', [(2, 16, 3), (2, 9, 4), (3, 20, 5), (3, 7, 2)])
def test_halo_exchange_and_gather_gloo(world, full_s0, halo):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, full_s0, halo, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]


def _blur_worker(rank, world, port, full_s0, sigma, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from oracle import conv as oconv
        g = torch.Generator().manual_seed(1)
        full = torch.randn(2, full_s0, 6, 7, 2, generator=g)
        z0, nz = nd.slab_bounds(full_s0, world, rank)
        blur = lambda t, s: torch.from_numpy(oconv.gaussian_blur(t.numpy(), s))      # noqa: E731
        part = nd.blur_slab(full[:, z0:z0 + nz].contiguous(), sigma, full_s0, blur_fn=blur)
        whole = blur(full, sigma if isinstance(sigma, list) else [sigma] * 3)
        q.put((rank, bool(torch.allclose(part, whole[:, z0:z0 + nz], rtol=0, atol=1e-6))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,full_s0,sigma', [(2, 12, 1.0), (3, 20, [1.5, 0.0, 0.7]), (2, 9, [0.0, 1.0, 1.0])])
def test_blur_slab_exchange_and_crop_gloo(world, full_s0, sigma):
    """z-slab GaussianBlur = halo exchange + blur + crop; checked with the oracle's blur on CPU tensors."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_blur_worker, args=(r, world, port, full_s0, sigma, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(30)
    assert all(ok for _, ok in res), res


def _oracle_warp_fn(full_s0):
    """SlabWarper's kernel hook on CPU tensors: the numpy oracle warps the FULL volume (planes outside the
    resident window are poisoned with NaN, so a read outside the window cannot go unnoticed) and the requested
    output planes are written into the caller's view."""
    import numpy as np
    from oracle import interp as ointerp

    def fn(vol_v, flow_v, out_v, src_z0, out_z0):
        B, n_src = vol_v.shape[0], vol_v.shape[1]
        rest, C = tuple(vol_v.shape[2:-1]), vol_v.shape[-1]
        src = np.full((B, full_s0) + rest + (C,), np.nan, dtype=np.float32)
        src[:, src_z0:src_z0 + n_src] = vol_v.numpy()
        flow = np.zeros((B, full_s0) + rest + (3,), dtype=np.float32)
        n_out = flow_v.shape[1]
        flow[:, out_z0:out_z0 + n_out] = flow_v.numpy()
        res = ointerp.spatial_transformer(src, flow)
        out_v.copy_(torch.from_numpy(np.ascontiguousarray(res[:, out_z0:out_z0 + n_out])))
    return fn


def _slab_warper_worker(rank, world, port, full_s0, amp, B, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from oracle import interp as ointerp
        g = torch.Generator().manual_seed(3)
        S = (full_s0, 6, 8)
        vol = torch.randn((B,) + S + (2,), generator=g)
        flow = (torch.rand((B,) + S + (3,), generator=g) * 2 - 1) * amp
        whole = torch.from_numpy(ointerp.spatial_transformer(vol.numpy(), flow.numpy()))
        z0, nz = nd.slab_bounds(full_s0, world, rank)
        halo = nd.agreed_halo(flow[:, z0:z0 + nz])
        plan = nd.SlabWarper(full_s0, halo, warp_fn=_oracle_warp_fn(full_s0))
        ok = plan.fits == nd.halo_fits(full_s0, world, halo)
        for step in range(2):                                   # the plan is reused: second call hits the cached buffers
            part = plan(vol[:, z0:z0 + nz].contiguous(), flow[:, z0:z0 + nz].contiguous())
            ok = ok and bool(torch.equal(part, whole[:, z0:z0 + nz])) and not bool(torch.isnan(part).any())
        # a producer may write its planes straight into the plan's buffer (no copy inside the call)
        view = plan.source_view(vol[:, z0:z0 + nz])
        view.copy_(vol[:, z0:z0 + nz])
        part = plan(view, flow[:, z0:z0 + nz].contiguous())
        ok = ok and bool(torch.equal(part, whole[:, z0:z0 + nz]))
        q.put((rank, bool(ok), plan.fits))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,full_s0,amp,B', [(2, 16, 2.5, 1), (3, 20, 1.5, 2), (2, 9, 2.0, 2), (3, 9, 4.5, 1)])
def test_slab_warper_overlapped_exchange_gloo(world, full_s0, amp, B):
    """SlabWarper on CPU tensors: interior launch from the rank's own planes, boundary launches after the
    neighbour exchange, thin slabs without an interior, and the all-gather branch when the halo exceeds a slab
    (last case) -- all ranks take the same branch and the pieces equal the whole-volume warp bit for bit."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_slab_warper_worker, args=(r, world, port, full_s0, amp, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(30)
    assert all(ok for _, ok, _ in res), res
    assert len({fits for _, _, fits in res}) == 1


def _halo_raise_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        z0, nz = nd.slab_bounds(10, world, rank)                # slabs of 4, 3, 3 planes
        slab = torch.zeros(1, nz, 2, 2, 1)
        try:
            nd.exchange_halo(slab, 4, 10)                       # fits rank 0's slab only
            q.put((rank, 'no error'))
        except ValueError:
            q.put((rank, 'raised'))
    finally:
        dist.destroy_process_group()


def test_halo_that_does_not_fit_raises_on_every_rank():
    """the fit decision comes from the slab table, so no rank is left blocking in a collective."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_halo_raise_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(3)]
    for p in procs:
        p.join(30)
    assert sorted(res) == [(0, 'raised'), (1, 'raised'), (2, 'raised')]
