"""
Multi-GPU tests (need >= 2 CUDA devices; skipped on a 1-GPU box).  NCCL process group, one
process per GPU.  z-slab warp (halo exchange and all-gather variants) and voxel-range
sharded Dice / CCE must equal the single-GPU result (bit-exact for the warp, 1e-6 for the
reductions).
"""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        import neurite_b200 as ne
        from neurite_b200 import dist as nd, utils
        dev = torch.device('cuda', rank)
        g = torch.Generator().manual_seed(0)
        S = (40, 24, 64)
        vol = torch.randn((2,) + S + (1,), generator=g).to(dev)
        ok = True
        for amp in (2.5, 7.0):                      # 7.0 > slab size at world 8 -> exercises the gather fallback
            flow = ((torch.rand((2,) + S + (3,), generator=g) * 2 - 1) * amp).to(dev)
            whole = utils._warp_batched(vol, flow)
            z0, nz = nd.slab_bounds(S[0], world, rank)
            for mode in ('auto', 'gather'):
                part = nd.warp_slab(vol[:, z0:z0 + nz].contiguous(), flow[:, z0:z0 + nz].contiguous(), S[0], mode=mode)
                ok = ok and torch.equal(part, whole[:, z0:z0 + nz])
        # Dice / CCE: voxel-range sharding + all-reduce of the partial sums
        L = 16
        lab = torch.randint(0, L, (2,) + S, generator=g)
        t = torch.nn.functional.one_hot(lab, L).float().to(dev)
        p = torch.softmax(torch.randn((2,) + S + (L,), generator=g), -1).to(dev)
        ref = ne.losses.Dice().loss(t, p)
        z0, nz = nd.slab_bounds(S[0], world, rank)
        sh = ne.losses.Dice(group=dist.group.WORLD).loss(t[:, z0:z0 + nz].contiguous(), p[:, z0:z0 + nz].contiguous())
        ok = ok and bool(torch.allclose(sh, ref, rtol=1e-6, atol=1e-7))
        cref = ne.losses.CategoricalCrossentropy().loss(t, p)
        csh = ne.losses.CategoricalCrossentropy(group=dist.group.WORLD).loss(t[:, z0:z0 + nz].contiguous(),
                                                                             p[:, z0:z0 + nz].contiguous())
        ok = ok and bool(torch.allclose(csh, cref, rtol=1e-6))
        # Resize: output slabs against a replicated source
        x = torch.randn((1, 10, 12, 16, 3), generator=g).to(dev)
        full = ne.layers.Resize(2)(x)
        z0, nz = nd.slab_bounds(full.shape[1], world, rank)
        ok = ok and torch.equal(nd.resize_slab(x, 2), full[:, z0:z0 + nz])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 4, 8])
def test_sharded_equals_single_gpu(world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip('needs %d GPUs' % world)
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]
