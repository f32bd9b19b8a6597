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
