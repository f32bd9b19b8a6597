#!/usr/bin/env python
"""
BASELINE.json configs[4]: UNet forward -> SpatialTransformer -> Dice on synthetic
160x192x224 volumes, one process per GPU.

    python examples/cfg5_unet_warp_dice.py --batch-per-gpu 1
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 \
        examples/cfg5_unet_warp_dice.py --batch-per-gpu 1 [--slab]

The UNet is context: the reference builds it from stock Keras conv /
pool / upsample layers (neurite/tf/models.py:88-246, conv_enc :1309, conv_dec :1445); here
the same topology is written with stock torch (cuDNN) layers only to drive the hot path.
The hot path itself -- the 16-channel warp and the Dice loss -- runs in neurite_b200's CUDA
kernels:

  default   batch-sharded: every rank pushes its own volumes through UNet -> warp -> Dice;
            the only collective is the all-reduce of the scalar mean loss.
  --slab    ONE volume's segmentation is warped and scored with the z axis split over the
            ranks (neurite_b200.dist): source halo exchange (or all-gather) for the warp,
            all-reduce of the [1,16,3] Dice partial sums -- the 8xB200 "z-slab shard with
            halo" variant of the config.
"""
import argparse
import os
import sys
import time

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neurite_b200 as ne  # noqa: E402
from neurite_b200 import dist as nd  # noqa: E402

SHAPE = (160, 192, 224)


class UNet(nn.Module):
    """ne.models.unet topology: nb_levels of [conv3-ELU] + maxpool, mirrored decoder with
    upsample + skip concat, 1x1x1 conv + softmax over nb_labels (models.py:88-246)."""

    def __init__(self, nb_features=16, nb_levels=4, nb_labels=16, feat_mult=2, in_ch=1):
        super().__init__()
        self.enc, self.dec = nn.ModuleList(), nn.ModuleList()
        ch, feats = in_ch, []
        for lvl in range(nb_levels):
            f = int(nb_features * feat_mult ** lvl)
            self.enc.append(nn.Conv3d(ch, f, 3, padding=1))
            feats.append(f)
            ch = f
        for lvl in range(nb_levels - 2, -1, -1):
            f = feats[lvl]
            self.dec.append(nn.Conv3d(ch + f, f, 3, padding=1))
            ch = f
        self.head = nn.Conv3d(ch, nb_labels, 1)

    def forward(self, x):                                   # x [B,1,D,H,W] (channels_last_3d memory)
        skips = []
        for i, conv in enumerate(self.enc):
            x = F.elu(conv(x))
            if i < len(self.enc) - 1:
                skips.append(x)
                x = F.max_pool3d(x, 2)
        for conv in self.dec:
            s = skips.pop()
            x = F.interpolate(x, size=s.shape[2:], mode='nearest')
            x = F.elu(conv(torch.cat([x, s], 1)))
        return torch.softmax(self.head(x), 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch-per-gpu', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--features', type=int, default=16)
    ap.add_argument('--levels', type=int, default=4)
    ap.add_argument('--shape', type=int, nargs=3, default=list(SHAPE))
    ap.add_argument('--slab', action='store_true')
    args = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        torch.distributed.init_process_group('nccl', device_id=dev)
    group = torch.distributed.group.WORLD if world > 1 else None
    S, L, B = tuple(args.shape), 16, args.batch_per_gpu
    torch.manual_seed(0)
    net = UNet(args.features, args.levels, L).to(dev).to(memory_format=torch.channels_last_3d).eval()
    g = torch.Generator(device=dev).manual_seed(100 + (0 if args.slab else rank))
    img = torch.randn((B, 1) + S, device=dev, generator=g).contiguous(memory_format=torch.channels_last_3d)
    flow = torch.rand((B,) + S + (3,), device=dev, generator=g) * 6 - 3
    target = F.one_hot(torch.randint(0, L, (B,) + S, device=dev, generator=g), L).float()
    warp = ne.layers.SpatialTransformer()

    def step():
        with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16):
            seg = net(img)                                   # [B,L,D,H,W], channels_last_3d memory
        seg = seg.float().permute(0, 2, 3, 4, 1).contiguous()   # channels-last view [B,D,H,W,L] (no copy if already NDHWC)
        if args.slab:
            z0, nz = nd.slab_bounds(S[0], world, rank)
            moved = nd.warp_slab(seg[:, z0:z0 + nz].contiguous(), flow[:, z0:z0 + nz].contiguous(), S[0], group=group) \
                if world > 1 else warp([seg, flow])
            tgt = target[:, z0:z0 + nz].contiguous() if world > 1 else target
            loss = ne.losses.Dice(group=group).mean_loss(tgt, moved.clamp_(0, 1))
        else:
            moved = warp([seg, flow])
            loss = ne.losses.Dice().mean_loss(target, moved.clamp_(0, 1))
            if world > 1:
                torch.distributed.all_reduce(loss)
                loss = loss / world
        return loss

    step()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    vols = args.steps * (B if args.slab else B * world)
    if rank == 0:
        print('cfg5 %s: %d GPU(s), %.1f ms/step, %.2f volumes/s, mean Dice loss %.5f'
              % ('z-slab' if args.slab else 'batch-sharded', world, float(ms) / args.steps,
                 vols / (float(ms) * 1e-3), float(loss)))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
