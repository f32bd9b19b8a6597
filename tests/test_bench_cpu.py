"""
CPU tests of bench.py's host-side pieces: the reference arm (`--impl reference`, the C/OpenMP port
of the reference on the host cores) prints one well-formed JSON line, and the bounded CPU-baseline
timers used by the `--op mi` / `--op blur` lines run.  (The GPU arm needs a device.)
"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '2', '--warmup', '1'],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'voxels/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['steps'] == 2 and d['gpu_launches'] == 0 and d['warmup'] >= 3
    assert d['config']['batch_per_gpu'] == 8 and 'MEDIAN' in d['cpu_baseline']['sample']
    assert d['e2e']['value'] == d['value'] and d['e2e']['h2d_bytes_per_step'] == 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
    assert '160x192x224' in d['metric'] and d['config']['volume'] == [160, 192, 224]


def test_bounded_cpu_baseline_timers():
    import bench
    from oracle import cport
    x = np.random.default_rng(0).uniform(0, 1, (1, 12, 14, 16, 1)).astype(np.float32)
    y = (x * x).astype(np.float32)
    r = bench.cpu_record(lambda: cport.mi_channelwise(x, y, nb_bins=16), x.size, 'voxels/s', 'tiny', budget_s=0.2, min_runs=5)
    assert r['value'] > 0 and r['unit'] == 'voxels/s' and r['kind'] == 'port' and r['cores'] >= 1 and 'median of' in r['sample']
    r = bench.cpu_record(lambda: cport.gaussian_blur(x, 1.0), x.size, 'voxels/s', 'tiny', budget_s=0.2, min_runs=5)
    assert r['value'] > 0
    # the statistic: median of the sorted run times, never fewer than min_runs runs
    calls = []
    med, n, best = bench.median_time(lambda: calls.append(1), budget_s=0.0, min_runs=7, warm=1)
    assert n == 7 and len(calls) == 8 and best <= med


def test_host_placement_helpers():
    import bench
    assert bench.parse_cpulist('0-3,8,10-11') == {0, 1, 2, 3, 8, 10, 11}
    assert bench.parse_cpulist('') == set()


def test_reference_arm_and_our_arm_share_one_config():
    """`same_config` was false because the reference arm's config lacked batch_per_gpu."""
    import argparse
    import bench
    a = argparse.Namespace(batch=8, flow='iid', method='linear')
    c = bench.warp_config(a, 1)
    assert c['batch_per_gpu'] == 8 and c['volume'] == [160, 192, 224] and 'configs[1]' in c['workload']


def test_gpu_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '1', '--warmup', '1'],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0 and 'no CUDA device' in (out.stderr + out.stdout)
