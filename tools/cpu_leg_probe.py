#!/usr/bin/env python
"""Dev tool (GPU box's host): why is the CPU leg bimodal?  Times the oracle's warp of ONE 160x192x224 volume with
different thread counts, binding and wait policies (one process per setting: libgomp reads the environment once)
and prints the quartiles of 30 runs next to the container's CPU quota and the host's load.

    python tools/cpu_leg_probe.py            # driver: spawns the settings
    python tools/cpu_leg_probe.py N          # worker: N threads, environment as given
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(n):
    import numpy as np
    from oracle import cport
    cport.build()
    cport.set_num_threads(n)
    S = (160, 192, 224)
    rng = np.random.default_rng(0)
    vol = cport.first_touch(rng.standard_normal((1,) + S + (1,)).astype(np.float32))
    flow = cport.first_touch(rng.uniform(-3, 3, (1,) + S + (3,)).astype(np.float32))
    out = cport.first_touch(np.zeros(vol.shape, np.float32))
    for _ in range(3):
        cport.warp(vol, flow, 'linear', out=out)
    ts = []
    for _ in range(30):
        t = time.time()
        cport.warp(vol, flow, 'linear', out=out)
        ts.append((time.time() - t) * 1e3)
    ts.sort()
    q = [ts[0], ts[len(ts) // 4], ts[len(ts) // 2], ts[3 * len(ts) // 4], ts[-1]]
    print('threads %3d bind=%-5s wait=%-7s  ms min/q1/med/q3/max = %s' % (
        cport.num_threads(), os.environ.get('OMP_PROC_BIND', '-'), os.environ.get('OMP_WAIT_POLICY', '-'),
        ' / '.join('%.1f' % v for v in q)), flush=True)


def main():
    if len(sys.argv) > 1:
        return worker(int(sys.argv[1]))
    from oracle import cport
    print('affinity cores', len(os.sched_getaffinity(0)), 'cgroup cpu limit', cport.cgroup_cpu_limit(), 'loadavg', os.getloadavg(), flush=True)
    for f in ('/sys/fs/cgroup/cpu.max', '/sys/fs/cgroup/cpu.stat', '/sys/fs/cgroup/cpu/cpu.stat', '/sys/fs/cgroup/cpuset.cpus.effective'):
        try:
            print(f, open(f).read().replace('\n', ' ')[:300], flush=True)
        except OSError:
            pass
    cores = len(os.sched_getaffinity(0))
    for bind, wait in (('close', 'active'), ('false', 'passive'), ('close', 'passive')):
        for n in sorted({cores, max(cores // 2, 1), max(cores // 4, 1), max(cores // 8, 1)}, reverse=True):
            env = dict(os.environ, OMP_PROC_BIND=bind, OMP_WAIT_POLICY=wait)
            if bind == 'close':
                env['OMP_PLACES'] = 'cores'
            else:
                env.pop('OMP_PLACES', None)
            subprocess.run([sys.executable, os.path.abspath(__file__), str(n)], env=env, timeout=300)
    print('loadavg after', os.getloadavg(), flush=True)


if __name__ == '__main__':
    main()
