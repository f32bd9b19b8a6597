"""
CPU model of the plane bookkeeping in resize3d_pair_kernel (neurite_b200/csrc/nrt_interp.cu, resize_pair_march): a
thread keeps the corner values of two source planes in two register sets P and Q, each tagged with the source plane it
holds; for an output plane with cell (i0, i1) the roles are (lo, hi) = (Q, P) if Q already holds i0, else (P, Q), and
only the sets whose tag is wrong are reloaded.  Pinned here with numpy / plain python:

  * lo always holds plane i0 and hi plane i1 (also at the clamped end, i0 == i1, and for down-sampling cells that skip),
  * marching up an up-sampling axis loads every source plane exactly once (no reload, no register copies; one extra
    load for the clamped last cell, whose two corners are the same plane),
  * the box origin rule of the kernel's issuing thread: [i0 of the tile's first output, i1 of its last output] covers
    every corner of the tile (the fp32 linspace is monotonic).
"""
import numpy as np
import pytest

F32 = np.float32


def cells(S, M):
    """(i0, i1) per output index, with the kernel's fp32 arithmetic (tf.linspace(0, S-1, M): endpoints exact)"""
    delta = F32(S - 1) / F32(M - 1) if M > 1 else F32(0)
    out = []
    for i in range(M):
        loc = F32(S - 1) if (i == M - 1 and M > 1) else F32(delta * F32(i))
        f0 = min(max(np.floor(loc), 0), S - 1)
        f1 = min(f0 + 1, S - 1)
        out.append((int(f0), int(f1)))
    return out


def march(seq):
    tag_p = tag_q = -1
    loads = 0
    for i0, i1 in seq:
        swapped = tag_q == i0
        need_p, need_q = (i1, i0) if swapped else (i0, i1)
        if tag_p != need_p:
            tag_p = need_p
            loads += 1
        if tag_q != need_q:
            tag_q = need_q
            loads += 1
        lo, hi = (tag_q, tag_p) if swapped else (tag_p, tag_q)
        assert (lo, hi) == (i0, i1)
    return loads


@pytest.mark.parametrize('S,M', [(80, 160), (96, 192), (7, 14), (9, 22), (5, 5), (33, 16), (3, 40), (1, 4), (6, 1), (2, 2),
                                 (40, 147), (112, 224)])
def test_plane_sets_always_hold_the_cell(S, M):
    seq = cells(S, M)
    loads = march(seq)
    planes = sorted({p for c in seq for p in c})
    if M >= S:                                           # up-sampling: consecutive cells, each plane loaded once
        assert len(planes) <= loads <= len(planes) + 1   # (+ 1: the clamped last cell holds plane S-1 in BOTH sets)
    else:
        assert loads <= 2 * len(seq)
    # a tile may start anywhere on the axis
    for z0 in range(0, M, 7):
        march(seq[z0:z0 + 32])


@pytest.mark.parametrize('S,M,T', [(80, 160, 32), (96, 192, 16), (112, 224, 32), (9, 22, 8), (40, 147, 32), (33, 16, 8)])
def test_box_origin_rule_covers_the_tile(S, M, T):
    seq = cells(S, M)
    i0s = [c[0] for c in seq]
    i1s = [c[1] for c in seq]
    assert i0s == sorted(i0s) and i1s == sorted(i1s)     # monotonic
    for t0 in range(0, M, T):
        t1 = min(t0 + T, M) - 1
        lo, hi = seq[t0][0], seq[t1][1]
        assert all(lo <= a and b <= hi for a, b in seq[t0:t1 + 1])
