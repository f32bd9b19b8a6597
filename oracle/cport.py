"""
ctypes binding of oracle/c/liboracle.so (C99 + OpenMP restatement; TEST INFRASTRUCTURE --
see oracle/__init__.py).  Used for full-size parity checks and as the multi-threaded CPU
baseline in bench.py.  tests/test_oracle_c.py pins it bit-exactly to oracle/interp.py.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, 'c', 'liboracle.so')
F32 = np.float32


def build(force=False):
    src = os.path.join(_HERE, 'c', 'nrt_oracle.c')
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(['make', '-s', '-C', os.path.join(_HERE, 'c')])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_cce_f32.restype = ctypes.c_double
        _lib.oracle_num_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _ints(v):
    return (ctypes.c_int * len(v))(*[int(x) for x in v])


def num_threads():
    return int(lib().oracle_num_threads())


def set_num_threads(n):
    """torchrun exports OMP_NUM_THREADS=1; the CPU baseline is asked to use every host core."""
    lib().oracle_set_num_threads(ctypes.c_int(int(n)))
    return num_threads()


def cgroup_cpu_limit():
    """CPUs' worth of run time per period the container may use (cgroup v2 cpu.max / v1 cfs quota), or None.
    More runnable threads than this get the whole group throttled until the next 100 ms period: bimodal timings
    (a 5 ms loop takes 100-200 ms most of the time)."""
    try:
        with open('/sys/fs/cgroup/cpu.max') as f:                      # v2: "<quota|max> <period>"
            q, per = f.read().split()[:2]
            return None if q == 'max' else float(q) / float(per)
    except (OSError, ValueError):
        pass
    try:
        with open('/sys/fs/cgroup/cpu/cpu.cfs_quota_us') as f:
            q = float(f.read())
        with open('/sys/fs/cgroup/cpu/cpu.cfs_period_us') as f:
            per = float(f.read())
        return None if q <= 0 else q / per
    except (OSError, ValueError):
        return None


def use_all_cores():
    """every core of the affinity mask, but no more threads than the container's CPU quota"""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lim = cgroup_cpu_limit()
    if lim is not None:
        n = max(1, min(n, int(lim)))
    return set_num_threads(n)


def first_touch(a):
    """Copy of `a` whose pages are first touched by the OpenMP threads with the static schedule of the compute
    loops (np.empty does not touch; the parallel copy does): NUMA-local inputs for the CPU baseline."""
    a = np.ascontiguousarray(a, dtype=F32)
    out = np.empty(a.shape, dtype=F32)
    lib().oracle_first_touch_copy_f32(_p(out), _p(a), ctypes.c_int64(a.size))
    return out


def warp(vol, flow, interp_method='linear', fill_value=None, out=None):
    """vol [B,*S,C], flow [B,*S,D] -> [B,*S,C] (written into `out` if given)"""
    vol = np.ascontiguousarray(vol, dtype=F32)
    flow = np.ascontiguousarray(flow, dtype=F32)
    D = flow.shape[-1]
    if out is None:
        out = np.empty(flow.shape[:-1] + (vol.shape[-1],), dtype=F32)
    assert out.dtype == F32 and out.flags.c_contiguous and out.shape == flow.shape[:-1] + (vol.shape[-1],)
    lib().oracle_warp_f32(_p(vol), _p(flow), _p(out), ctypes.c_int(vol.shape[0]), _ints(vol.shape[1:-1]),
                          ctypes.c_int(D), ctypes.c_int(vol.shape[-1]),
                          ctypes.c_int(0 if interp_method == 'linear' else 1),
                          ctypes.c_int(fill_value is not None), ctypes.c_float(fill_value or 0.0))
    return out


def interpn(vol, loc, interp_method='linear', fill_value=None):
    """vol [*S, C], loc [*O, D] -> [*O, C]"""
    vol = np.ascontiguousarray(vol, dtype=F32)
    loc = np.ascontiguousarray(loc, dtype=F32)
    D = loc.shape[-1]
    out = np.empty(loc.shape[:-1] + (vol.shape[-1],), dtype=F32)
    lib().oracle_interpn_f32(_p(vol), _ints(vol.shape[:-1]), ctypes.c_int(D), ctypes.c_int(vol.shape[-1]), _p(loc),
                             ctypes.c_int64(loc.size // D), ctypes.c_int(0 if interp_method == 'linear' else 1),
                             ctypes.c_int(fill_value is not None), ctypes.c_float(fill_value or 0.0), _p(out))
    return out


def dice_sums(t, p):
    t = np.ascontiguousarray(t, dtype=F32)
    p = np.ascontiguousarray(p, dtype=F32)
    B, L = t.shape[0], t.shape[-1]
    V = t.size // (B * L)
    sums = np.empty((B, L, 3), dtype=F32)
    lib().oracle_dice_sums_f32(_p(t), _p(p), ctypes.c_int(B), ctypes.c_int64(V), ctypes.c_int(L), _p(sums))
    return sums


def dice(t, p):
    s = dice_sums(t, p)
    top = F32(2) * s[..., 0]
    bottom = s[..., 1] + s[..., 2]
    out = np.zeros_like(top)
    np.divide(top, bottom, out=out, where=bottom != 0)
    return out


def cce(t, p, label_weights=None):
    t = np.ascontiguousarray(t, dtype=F32)
    p = np.ascontiguousarray(p, dtype=F32)
    C = t.shape[-1]
    lw = None if label_weights is None else np.ascontiguousarray(label_weights, dtype=F32)
    n = t.size // C
    tot = lib().oracle_cce_f32(_p(t), _p(p), None if lw is None else _p(lw), ctypes.c_int64(n), ctypes.c_int(C))
    return F32(tot / n)


def lc3d(x, kernel, bias, kernel_size, strides=(1, 1, 1)):
    x = np.ascontiguousarray(x, dtype=F32)
    kernel = np.ascontiguousarray(kernel, dtype=F32)
    B, Cin, Cout = x.shape[0], x.shape[-1], kernel.shape[-1]
    O = [(x.shape[1 + d] - kernel_size[d]) // strides[d] + 1 for d in range(3)]
    out = np.empty((B, O[0], O[1], O[2], Cout), dtype=F32)
    b = None if bias is None else np.ascontiguousarray(bias, dtype=F32)
    lib().oracle_lc3d_f32(_p(x), _p(kernel), None if b is None else _p(b), _p(out), ctypes.c_int(B),
                          _ints(x.shape[1:4]), ctypes.c_int(Cin), ctypes.c_int(Cout), _ints(kernel_size), _ints(strides))
    return out


def mi_channelwise(x, y, nb_bins=16, alpha=None, min_clip=-np.inf, max_clip=np.inf, bin_centers=None):
    """MutualInformation.channelwise on [B, ..., C] intensity tensors -> [B, C] (metrics.py:185-292)."""
    from . import mi as _mi
    x = np.ascontiguousarray(x, dtype=F32)
    y = np.ascontiguousarray(y, dtype=F32)
    B, C = x.shape[0], x.shape[-1]
    V = x.size // (B * C)
    if bin_centers is not None:
        cen = np.ascontiguousarray(bin_centers, dtype=F32)
        nb_bins = cen.shape[0]
    if alpha is None:
        alpha = _mi.default_alpha(nb_bins, bin_centers)
    out = np.empty((B, C), dtype=F32)
    lib().oracle_mi_channelwise_f32(_p(x), _p(y), ctypes.c_int(B), ctypes.c_int64(V), ctypes.c_int(C),
                                    ctypes.c_int(int(nb_bins)), ctypes.c_float(float(alpha)),
                                    ctypes.c_float(float(min_clip)), ctypes.c_float(float(max_clip)),
                                    _p(cen) if bin_centers is not None else None,
                                    _p(cen) if bin_centers is not None else None, _p(out))
    return out


def gaussian_blur(x, sigma, out=None):
    """layers.GaussianBlur(sigma)(x) for x [B, *space, C] (non-random): one C pass per blurred axis."""
    from . import conv as _conv
    x = np.ascontiguousarray(x, dtype=F32)
    nd = x.ndim - 2
    sig = np.ravel(sigma).tolist()
    sig = sig * nd if len(sig) == 1 else sig
    if not any(s > 0 for s in sig):
        return x
    ks = _conv.gaussian_kernel(sig, separate=True)
    ks = ks if isinstance(ks, list) else [ks]
    cur = x
    for ax, k in enumerate(ks):
        K = len(k)
        if K == 1 and k[0] == 1:
            continue
        shp = cur.shape
        outer = int(np.prod(shp[:ax + 1], dtype=np.int64))
        L = shp[ax + 1]
        inner = int(np.prod(shp[ax + 2:], dtype=np.int64))
        nxt = np.empty_like(cur)
        k = np.ascontiguousarray(k, dtype=F32)
        lib().oracle_sepconv_axis_f32(_p(cur), _p(nxt), ctypes.c_int64(outer), ctypes.c_int64(L), ctypes.c_int64(inner),
                                      _p(k), ctypes.c_int(K), ctypes.c_int(1), ctypes.c_int(1),
                                      ctypes.c_int((K - 1) // 2), ctypes.c_int64(L))
        cur = nxt
    return cur
