#!/usr/bin/env python
"""
bench.py -- BASELINE.json's metric: voxels/s warped (SpatialTransformer / interpn linear) on
160x192x224 fp32 volumes, with the achieved fraction of the HBM roofline.

    python bench.py --gpus 1 --steps K --warmup W                 # our CUDA path
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                           # CPU port of the reference, host cores
    python bench.py --op dice|cce|lc3d|resize ...                  # the other configs (extra lines)

A "step" is one pass of the hot path over one batch: `--batch` (default 8) independent
160x192x224x1 volumes with a random dense 3-channel flow U(-3,3) (configs[1] of
BASELINE.json), one kernel launch.  The batch's working set is 1.1 GB, far
larger than the 126 MB L2, so no flush is needed between iterations.  At N GPUs every rank
warps its own batch (weak scaling, no data-path collective); `value` is whole-job
voxels/s = N * batch * V * K / max-over-ranks device time (CUDA events, barrier + sync on
both sides).

JSON keys beyond the base contract:
  roofline     achieved = 20 B/voxel (12 flow + 4 source-once + 4 store)
               * voxels per launch / launch time; peak = MEASURED_PEAKS.json hbm_gbs.
  e2e          the same metric through the public API with HOST (pinned) buffers: H2D of
               vol+flow and D2H of the result inside the timed region, every step.
  cpu_baseline the oracle's C/OpenMP port on the host cores, bounded sample (rank 0, N=1).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPE = (160, 192, 224)
V = SHAPE[0] * SHAPE[1] * SHAPE[2]


def measured_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:                                       # noqa: BLE001
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def ncu_traffic(op):
    """dram bytes per launch from the committed ncu summary, if present (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            return json.load(f).get(op)
    except Exception:                                       # noqa: BLE001
        return None


class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the benchmark runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:                                   # noqa: BLE001
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {'hw_slowdown': 0x8, 'sw_power_cap': 0x4, 'sw_thermal_slowdown': 0x20,
                 'hw_thermal_slowdown': 0x40, 'hw_power_brake': 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:                               # noqa: BLE001
                pass
            time.sleep(0.01)

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()
        s = sorted(self.samples)
        return {'sm_mhz': s[len(s) // 2] if s else None, 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons), 'samples': len(s)}


def dist_setup(n_gpus):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    return world, rank, local


def timed_region(fn, steps, warmup, world, min_preheat_s=0.3):
    """W warm-up steps, then exactly K steps between CUDA events, barrier + sync both sides,
    max over ranks.  A short pre-heat (not counted) lets the clocks settle and gives the
    sampler something to see."""
    import torch
    import torch.distributed as dist
    t0 = time.time()
    while time.time() - t0 < min_preheat_s:
        fn()
        torch.cuda.synchronize()
    for _ in range(max(warmup, 3)):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


# ---------------------------------------------------------------------------------------
def bench_warp(args):
    import torch
    import neurite_b200 as ne
    world, rank, local = dist_setup(args.gpus)
    dev = torch.device('cuda', local)
    B = args.batch
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    vol = torch.randn((B,) + SHAPE + (1,), device=dev, generator=g)
    flow = torch.rand((B,) + SHAPE + (3,), device=dev, generator=g) * 6 - 3
    if args.flow == 'smooth':
        # low-frequency field, max |u| = 8 voxels (SURVEY.md 8d secondary run)
        coarse = torch.randn((B, 3, 10, 12, 14), device=dev, generator=g)
        flow = torch.nn.functional.interpolate(coarse, size=SHAPE, mode='trilinear', align_corners=True)
        flow = (flow / flow.abs().amax() * 8).permute(0, 2, 3, 4, 1).contiguous()
    st = ne.layers.SpatialTransformer(interp_method=args.method, fill_value=None, halo=args.halo)
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed_region(lambda: st([vol, flow]), args.steps, args.warmup, world)
    clocks = sampler.stop()
    vox_per_step = world * B * V
    value = vox_per_step * args.steps / (ms * 1e-3)
    peak, peak_src = measured_peak()
    bytes_per_launch = 20.0 * B * V
    achieved = bytes_per_launch * args.steps / (ms * 1e-3) / 1e9           # per GPU (one launch per step per rank)

    # ---- end to end through the public API with host buffers
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    h_vol = torch.empty(vol.shape, dtype=torch.float32).pin_memory().copy_(vol.cpu())
    h_flow = torch.empty(flow.shape, dtype=torch.float32).pin_memory().copy_(flow.cpu())
    h_out = torch.empty(vol.shape, dtype=torch.float32).pin_memory()

    def e2e_step():
        # public host-facing call: pinned host tensors in, pinned host tensor out; returns
        # only when the result is in h_out (H2D of vol+flow and D2H of the result inside)
        st.call_host([h_vol, h_flow], out=h_out)
    ms_e2e = timed_region(e2e_step, e2e_steps, 2, world, min_preheat_s=0.0)
    e2e_value = vox_per_step * e2e_steps / (ms_e2e * 1e-3)

    line = {
        'metric': 'voxels/s warped (SpatialTransformer / interpn %s), 160x192x224 fp32' % args.method,
        'value': value, 'unit': 'voxels/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'SpatialTransformer warp of 160x192x224x1 fp32 volumes, random dense 3-ch flow '
                               '(%s), batch %d per GPU per step (BASELINE.json configs[1])'
                               % ('U(-3,3) i.i.d.' if args.flow == 'iid' else 'smooth, max|u|=8', B),
                   'batch_per_gpu': B, 'volume': list(SHAPE), 'channels': 1, 'interp_method': args.method,
                   'parallelism': 'batch-sharded x%d, no collective' % world,
                   'l2': 'working set %.2f GB per step > 126 MB L2 (no flush needed)' % (bytes_per_launch / 1e9)},
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     'traffic': ncu_traffic('warp'), 'peak_source': peak_src,
                     'bytes_model': '20 B/voxel = 12 flow + 4 source (each voxel once) + 4 store',
                     'kernel': 'warp3d_tile_kernel', 'per': 'GPU'},
        'e2e': {'value': e2e_value, 'unit': 'voxels/s', 'h2d_bytes_per_step': int(h_vol.numel() + h_flow.numel()) * 4,
                'd2h_bytes_per_step': int(h_out.numel()) * 4, 'steps': e2e_steps, 'ms_per_step': ms_e2e / e2e_steps},
        'gpu_launches': args.steps,
        'clocks': clocks,
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_baseline_warp(args.method, budget_s=8.0)
            if not args.no_numpy_baseline:
                line['cpu_baseline_numpy'] = cpu_baseline_warp_numpy(args.method)
        print(json.dumps(line), flush=True)
    finish(world)


def synth_host_volume(seed=0):
    import numpy as np
    vol = np.random.default_rng(seed).standard_normal((1,) + SHAPE + (1,)).astype(np.float32)
    flow = np.random.default_rng(seed + 1).uniform(-3, 3, (1,) + SHAPE + (3,)).astype(np.float32)
    return vol, flow


def cpu_baseline_warp(method, budget_s=8.0, vol=None, flow=None):
    """The oracle's C/OpenMP port (same arithmetic as the reference, fused, all host threads)."""
    from oracle import cport
    cport.build()
    cport.use_all_cores()
    if vol is None:
        vol, flow = synth_host_volume()
    import numpy as np
    out = np.zeros(vol.shape, dtype=np.float32)            # pre-faulted output, reused like a real pipeline would
    cport.warp(vol, flow, method, out=out)                 # warm-up
    best, n, t_all = None, 0, time.time()
    while n < 3 or (time.time() - t_all < budget_s and n < 50):
        t = time.time()
        cport.warp(vol, flow, method, out=out)
        dt = time.time() - t
        best = dt if best is None else min(best, dt)
        n += 1
    return {'value': V / best, 'unit': 'voxels/s', 'cores': cport.num_threads(), 'kind': 'port',
            'sample': 'oracle/c (C99+OpenMP restatement of interpn, -ffp-contract=off), one 160x192x224 volume, '
                      'best of %d runs' % n}


def cpu_baseline_warp_numpy(method):
    """op-for-op numpy restatement (mirrors the reference's unfused TF op sequence), 1 thread."""
    from oracle import interp
    vol, flow = synth_host_volume()
    t = time.time()
    interp.spatial_transformer(vol, flow, method)
    dt = time.time() - t
    return {'value': V / dt, 'unit': 'voxels/s', 'cores': 1, 'kind': 'port',
            'sample': 'oracle/interp.py (numpy, op-for-op like the reference TF graph), one 160x192x224 volume, 1 run'}


def bench_reference(args):
    """--impl reference: the reference's CPU path (its C/OpenMP port; TensorFlow is not
    installable here) on the host cores.  Rank 0 only."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from oracle import cport
    cport.build()
    cport.use_all_cores()
    import numpy as np
    vol, flow = synth_host_volume()
    out = np.zeros(vol.shape, dtype=np.float32)            # pre-faulted output buffer, reused every step
    for _ in range(max(1, min(args.warmup, 3))):
        cport.warp(vol, flow, args.method, out=out)
    steps = max(1, min(args.steps, 200))
    t = time.time()
    for _ in range(steps):
        cport.warp(vol, flow, args.method, out=out)
    dt = time.time() - t
    value = V * steps / dt
    sample = ('each step = ONE 160x192x224 volume (bounded sample of the batch-%d step), oracle/c C99+OpenMP port, '
              '%d threads' % (args.batch, cport.num_threads()))
    print(json.dumps({
        'impl': 'reference',
        'metric': 'voxels/s warped (SpatialTransformer / interpn %s), 160x192x224 fp32' % args.method,
        'value': value, 'unit': 'voxels/s', 'n_gpus': int(os.environ.get('WORLD_SIZE', '1')), 'steps': steps,
        'warmup': max(1, min(args.warmup, 3)), 'ms_per_step': dt / steps * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'SpatialTransformer warp of 160x192x224x1 fp32 volumes, random dense 3-ch flow '
                               'U(-3,3) i.i.d. (BASELINE.json configs[1])', 'volume': list(SHAPE), 'channels': 1,
                   'interp_method': args.method},
        'cpu_baseline': {'value': value, 'unit': 'voxels/s', 'cores': cport.num_threads(), 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': value, 'unit': 'voxels/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }), flush=True)


# ---------------------------------------------------------------------------------------
# the other configs (extra lines; same timing discipline)
# ---------------------------------------------------------------------------------------
def bench_dice(args, cce=False):
    import torch
    import neurite_b200 as ne
    world, rank, local = dist_setup(args.gpus)
    dev = torch.device('cuda', local)
    B, L = 4, 16
    # cfg 3: batch 4 < 8 GPUs -> shard the voxel range of every batch item across ranks (strong scaling)
    from neurite_b200.dist import slab_bounds
    z0, nz = slab_bounds(SHAPE[0], world, rank)
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    lab = torch.randint(0, L, (B, nz) + SHAPE[1:], device=dev, generator=g)
    t = torch.nn.functional.one_hot(lab, L).float()
    p = torch.softmax(torch.randn((B, nz) + SHAPE[1:] + (L,), device=dev, generator=g), -1)
    group = torch.distributed.group.WORLD if world > 1 else None
    if cce:
        op = ne.losses.CategoricalCrossentropy(group=group)
        fn = lambda: op.loss(t, p)                          # noqa: E731
    else:
        op = ne.losses.Dice(group=group)
        fn = lambda: op.loss(t, p)                          # noqa: E731
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed_region(fn, args.steps, args.warmup, world)
    clocks = sampler.stop()
    elems = B * V * L
    peak, peak_src = measured_peak()
    achieved = 8.0 * (B * nz * SHAPE[1] * SHAPE[2] * L) * args.steps / (ms * 1e-3) / 1e9
    if rank == 0:
        print(json.dumps({
            'metric': '(voxel,label) elements/s, %s on 16-label one-hot 160x192x224, batch 4' % ('CCE' if cce else 'Dice loss'),
            'value': elems * args.steps / (ms * 1e-3), 'unit': 'elements/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'BASELINE.json configs[2]: %s, y_true one-hot / y_pred softmax [4,160,192,224,16], '
                                   'voxel range sharded over %d GPU(s) + all-reduce of [4,16,3] partial sums'
                                   % ('CategoricalCrossentropy' if cce else 'Dice().loss', world),
                       'l2': '3.5 GB read per step > L2'},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': ncu_traffic('cce' if cce else 'dice'), 'peak_source': peak_src,
                         'bytes_model': '8 B per (voxel,label)', 'per': 'GPU'},
            'gpu_launches': args.steps * (2 if cce else 3), 'clocks': clocks}), flush=True)
    finish(world)


def bench_lc3d(args):
    import torch
    from neurite_b200.layers import local_conv3d
    world, rank, local = dist_setup(args.gpus)
    dev = torch.device('cuda', local)
    B = args.lc_batch
    I, Cin, Cout = 64, 16, 16
    O = I - 2
    P, F = O ** 3, 27 * Cin
    g = torch.Generator(device=dev).manual_seed(3)
    x = torch.randn((B, I, I, I, Cin), device=dev, generator=g)
    lim = (6.0 / (F + Cout)) ** 0.5
    kernel = (torch.rand((P, F, Cout), device=dev, generator=g) * 2 - 1) * lim
    bias = torch.randn((O, O, O, Cout), device=dev, generator=g)
    fn = lambda: local_conv3d(x, kernel, bias, (3, 3, 3), (1, 1, 1), (O, O, O))     # noqa: E731
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed_region(fn, args.steps, args.warmup, world)
    clocks = sampler.stop()
    peak, peak_src = measured_peak()
    nbytes = 4.0 * (P * F * Cout + B * I ** 3 * Cin + B * P * Cout + P * Cout)
    achieved = nbytes * args.steps / (ms * 1e-3) / 1e9
    if rank == 0:
        print(json.dumps({
            'metric': 'output positions/s, LocallyConnected3D 3^3 16->16 on 64^3, batch %d' % B,
            'value': P * B * args.steps / (ms * 1e-3), 'unit': 'positions/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'BASELINE.json configs[3]: LocallyConnected3D 3x3x3, 16->16, input [%d,64,64,64,16], '
                                   'kernel [238328,432,16] = 6.59 GB streamed once per step (> L2)' % B},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': ncu_traffic('lc3d'), 'peak_source': peak_src,
                         'bytes_model': '4*(P*F*Cout + B*in + B*P*Cout + P*Cout)', 'per': 'GPU'},
            'gpu_launches': args.steps, 'clocks': clocks}), flush=True)
    finish(world)


def bench_resize(args):
    import torch
    import neurite_b200 as ne
    world, rank, local = dist_setup(args.gpus)
    dev = torch.device('cuda', local)
    B = args.batch
    x = torch.randn((B, 80, 96, 112, 3), device=dev)
    lay = ne.layers.Resize(2)
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed_region(lambda: lay(x), args.steps, args.warmup, world)
    clocks = sampler.stop()
    peak, peak_src = measured_peak()
    nbytes = 4.0 * 3 * B * V * (1 + 1 / 8)
    achieved = nbytes * args.steps / (ms * 1e-3) / 1e9
    if rank == 0:
        print(json.dumps({
            'metric': 'output voxels/s, Resize zoom 2 of a half-resolution 3-ch flow to 160x192x224',
            'value': world * B * V * args.steps / (ms * 1e-3), 'unit': 'voxels/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'Resize(2) on [%d,80,96,112,3] (reference models.py:803-804)' % B},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': ncu_traffic('resize'), 'peak_source': peak_src, 'bytes_model': '4C/z^3 + 4C per output voxel', 'per': 'GPU'},
            'gpu_launches': args.steps, 'clocks': clocks}), flush=True)
    finish(world)


def cpu_baseline_timed(fn, nunits, unit, sample, budget_s=8.0):
    """best-of-n wall time of a C/OpenMP oracle call on the host cores (bounded: ~budget_s seconds)."""
    from oracle import cport
    cport.build()
    cport.use_all_cores()
    fn()                                                   # warm-up
    best, n, t_all = None, 0, time.time()
    while n < 2 or (time.time() - t_all < budget_s and n < 20):
        t = time.time()
        fn()
        dt = time.time() - t
        best = dt if best is None else min(best, dt)
        n += 1
    return {'value': nunits / best, 'unit': unit, 'cores': cport.num_threads(), 'kind': 'port',
            'sample': sample + ', best of %d runs' % n}


def bench_mi(args, segs=False):
    """MutualInformation: `mi` = volumes() on B pairs of 160x192x224 volumes (soft quantisation
    fused, 8 B/voxel); `mi_segs` = segs() on two [2,160,192,224,16] probability maps (128 B/voxel)."""
    import torch
    import neurite_b200 as ne
    world, rank, local = dist_setup(args.gpus)
    dev = torch.device('cuda', local)
    m = ne.metrics.MutualInformation(nb_bins=16)
    if segs:
        B = 2
        x = torch.softmax(torch.randn((B,) + SHAPE + (16,), device=dev), -1)
        y = torch.softmax(torch.randn((B,) + SHAPE + (16,), device=dev), -1)
        fn, per_voxel, kern = (lambda: m.segs(x, y)), 128.0, 'mi_hist_mma_kernel<1,2,maps,maps>'
    else:
        B = args.batch
        x = torch.rand((B,) + SHAPE + (1,), device=dev)
        y = (0.7 * x * x + 0.1 + 0.1 * torch.rand_like(x)).clamp_(0, 1)
        fn, per_voxel, kern = (lambda: m.volumes(x, y)), 8.0, 'mi_hist_mma_kernel<1,2,quant,quant>'
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed_region(fn, args.steps, args.warmup, world)
    clocks = sampler.stop()
    peak, peak_src = measured_peak()
    achieved = per_voxel * B * V * args.steps / (ms * 1e-3) / 1e9
    if rank == 0:
        line = ({
            'metric': 'voxels/s, MutualInformation.%s (16 bins), 160x192x224 fp32' % ('segs' if segs else 'volumes'),
            'value': world * B * V * args.steps / (ms * 1e-3), 'unit': 'voxels/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32 (3xTF32 tensor-core contraction)', 'data': 'synthetic',
            'config': {'workload': 'MutualInformation(nb_bins=16).%s on %d x 160x192x224 (reference metrics.py:41-336); '
                                   'min/max + histogram + combine + finalise kernels per step'
                                   % ('segs, 16 labels' if segs else 'volumes', B)},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': ncu_traffic('mi_segs' if segs else 'mi'), 'peak_source': peak_src,
                         'bytes_model': '%d B/voxel (two fp32 %s read once; the quantised [V,16] maps never exist)'
                                        % (per_voxel, 'maps' if segs else 'volumes'),
                         'kernel': kern, 'per': 'GPU',
                         'note': '' if segs else 'volumes(): bound by issue slots / MUFU (32 exp per voxel pair), not HBM'},
            'gpu_launches': args.steps * (4 if segs else 10), 'clocks': clocks})
        if world == 1 and not segs and not args.no_cpu_baseline:
            import numpy as np
            from oracle import cport
            xs = np.random.default_rng(0).uniform(0, 1, (1,) + SHAPE + (1,)).astype(np.float32)
            ys = np.clip(0.7 * xs * xs + 0.1 + 0.1 * np.random.default_rng(1).uniform(0, 1, xs.shape), 0, 1).astype(np.float32)
            line['cpu_baseline'] = cpu_baseline_timed(
                lambda: cport.mi_channelwise(xs, ys, nb_bins=16), V, 'voxels/s',
                'oracle/c oracle_mi_channelwise_f32 (C99+OpenMP restatement of metrics.py:185-292 with soft_quantize '
                'fused), ONE 160x192x224 volume pair')
        print(json.dumps(line), flush=True)
    finish(world)


def bench_blur(args):
    """GaussianBlur(sigma=1) (7 taps per axis) of B single-channel 160x192x224 volumes: three separable passes
    (NRT_BLUR_FUSED=1: the single fused kernel, measured slower)."""
    import torch
    import neurite_b200 as ne
    world, rank, local = dist_setup(args.gpus)
    dev = torch.device('cuda', local)
    B = args.batch
    x = torch.randn((B,) + SHAPE + (1,), device=dev)
    lay = ne.layers.GaussianBlur(sigma=args.sigma)
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed_region(lambda: lay(x), args.steps, args.warmup, world)
    clocks = sampler.stop()
    peak, peak_src = measured_peak()
    achieved = 8.0 * B * V * args.steps / (ms * 1e-3) / 1e9
    fused = os.environ.get('NRT_BLUR_FUSED', '0') == '1' and round(args.sigma * 3) * 2 + 1 <= 15
    if rank == 0:
        line = ({
            'metric': 'voxels/s, GaussianBlur(sigma=%g), 160x192x224 fp32' % args.sigma,
            'value': world * B * V * args.steps / (ms * 1e-3), 'unit': 'voxels/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'GaussianBlur(sigma=%g) on [%d,160,192,224,1] (reference layers.py:251-364): %s'
                                   % (args.sigma, B, 'one fused kernel' if fused else 'three separable passes')},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': ncu_traffic('blur_fused' if fused else 'blur'), 'peak_source': peak_src,
                         'bytes_model': '8 B/voxel for the whole blur (read once, write once)'
                                        + ('' if fused else '; the three-pass path moves 24 B/voxel, so 0.33 is its ceiling'),
                         'kernel': 'blur3d_fused_kernel' if fused else 'sepconv_col4_kernel x2 + sepconv_row_kernel',
                         'per': 'GPU'},
            'gpu_launches': args.steps * (1 if fused else 3), 'clocks': clocks})
        if world == 1 and not args.no_cpu_baseline:
            import numpy as np
            from oracle import cport
            xs = np.random.default_rng(0).standard_normal((1,) + SHAPE + (1,)).astype(np.float32)
            line['cpu_baseline'] = cpu_baseline_timed(
                lambda: cport.gaussian_blur(xs, args.sigma), V, 'voxels/s',
                'oracle/c oracle_sepconv_axis_f32 x3 (C99+OpenMP restatement of utils.py:665-751), ONE 160x192x224 volume')
        print(json.dumps(line), flush=True)
    finish(world)


def finish(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--op', default='warp', choices=['warp', 'dice', 'cce', 'lc3d', 'resize', 'mi', 'mi_segs', 'blur'])
    ap.add_argument('--sigma', type=float, default=1.0)
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--lc-batch', type=int, default=1)
    ap.add_argument('--method', default='linear', choices=['linear', 'nearest'])
    ap.add_argument('--flow', default='iid', choices=['iid', 'smooth'])
    ap.add_argument('--halo', type=int, default=0)
    ap.add_argument('--e2e-steps', type=int, default=10)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-numpy-baseline', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        return bench_reference(args)
    import torch
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the product path has no CPU fallback '
                         '(use --impl reference for the CPU port of the reference)')
    {'warp': bench_warp, 'dice': bench_dice, 'cce': lambda a: bench_dice(a, cce=True), 'lc3d': bench_lc3d,
     'resize': bench_resize, 'mi': bench_mi, 'mi_segs': lambda a: bench_mi(a, segs=True),
     'blur': bench_blur}[args.op](args)


if __name__ == '__main__':
    main()
