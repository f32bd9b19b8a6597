"""
CPU property tests (hypothesis) of the oracle -- the size-independent invariants the GPU
parity tests rely on at BASELINE.json's full sizes: identity warp is exact,
zoom endpoints hit the corner voxels, out-of-range samples clamp to the edge, fill_value uses
strict bounds, nearest rounds half to even, Dice(x,x)=1 / symmetric / 0 for disjoint labels,
CCE of a one-hot prediction = -log(1-1e-7).
"""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import cport, interp, metrics

F32 = np.float32
shapes3 = st.tuples(st.integers(2, 7), st.integers(2, 7), st.integers(2, 9))


@settings(max_examples=25, deadline=None)
@given(shapes3, st.integers(1, 3), st.integers(0, 2 ** 31 - 1))
def test_identity_warp_and_integer_shifts_are_exact(shape, C, seed):
    rng = np.random.default_rng(seed)
    vol = rng.standard_normal((1,) + shape + (C,)).astype(F32)
    zero = np.zeros((1,) + shape + (3,), F32)
    np.testing.assert_array_equal(interp.spatial_transformer(vol, zero), vol)
    np.testing.assert_array_equal(cport.warp(vol, zero), vol)
    shift = zero.copy()
    shift[..., 2] = 1.0                                      # whole-voxel shift along the last axis, edge clamped
    out = interp.spatial_transformer(vol, shift)
    np.testing.assert_array_equal(out[:, :, :, :-1], vol[:, :, :, 1:])
    np.testing.assert_array_equal(out[:, :, :, -1], vol[:, :, :, -1])


@settings(max_examples=25, deadline=None)
@given(shapes3, st.integers(0, 2 ** 31 - 1))
def test_out_of_range_clamps_and_fill_is_strict(shape, seed):
    rng = np.random.default_rng(seed)
    vol = rng.standard_normal(shape).astype(F32)
    mx = np.array(shape, F32) - 1
    loc = np.stack([-rng.uniform(0.1, 50, 8).astype(F32), mx[1] + rng.uniform(0.1, 50, 8).astype(F32),
                    np.full(8, mx[2] / 2, F32)], -1)
    out = interp.interpn(vol, loc)
    ref = interp.interpn(vol, np.stack([np.zeros(8, F32), np.full(8, mx[1]), loc[:, 2]], -1))
    np.testing.assert_array_equal(out, ref)                 # clamped to the edge voxel line
    np.testing.assert_array_equal(interp.interpn(vol, loc, fill_value=7.5), np.full(8, 7.5, F32))
    edge = np.array([[0, 0, 0], mx], F32)                   # loc == 0 and loc == max are in bounds (strict < / >)
    np.testing.assert_array_equal(interp.interpn(vol, edge, fill_value=7.5), [vol[0, 0, 0], vol[-1, -1, -1]])


@settings(max_examples=20, deadline=None)
@given(st.integers(2, 9), st.integers(2, 6), st.integers(0, 2 ** 31 - 1))
def test_zoom_endpoints_and_nearest_ties(n, z, seed):
    rng = np.random.default_rng(seed)
    vol = rng.standard_normal((n, n, 1)).astype(F32)
    out = interp.resize(vol, z)
    assert out.shape == (n * z, n * z, 1)
    for a in (0, -1):
        for b in (0, -1):
            assert out[a, b, 0] == vol[a, b, 0]             # linspace endpoints are exact
    line = np.arange(n + 1, dtype=F32)
    ties = (np.arange(n, dtype=F32) + 0.5)[:, None]
    r = interp.interpn(line, ties, 'nearest')
    np.testing.assert_array_equal(r, np.rint(ties[:, 0]))    # half to even: 0.5->0, 1.5->2, 2.5->2 ...


@settings(max_examples=20, deadline=None)
@given(st.integers(1, 3), st.integers(2, 9), st.integers(0, 2 ** 31 - 1))
def test_dice_and_cce_invariants(B, L, seed):
    rng = np.random.default_rng(seed)
    lab = rng.integers(0, L, (B, 5, 6))
    t = np.eye(L, dtype=F32)[lab]
    p = rng.uniform(0, 1, t.shape).astype(F32)
    d = metrics.Dice()
    present = np.stack([[np.any(lab[b] == l) for l in range(L)] for b in range(B)])
    np.testing.assert_array_equal(d.dice(t, t), present.astype(F32))
    np.testing.assert_allclose(d.dice(t, p), d.dice(p, t), rtol=1e-6)
    other = np.eye(L, dtype=F32)[(lab + 1) % L]
    if L > 1:
        np.testing.assert_array_equal(d.dice(t, other) * (~present | True), d.dice(t, other))
        assert np.all(d.dice(t, other)[present & np.stack([[not np.any((lab[b] + 1) % L == l) for l in range(L)] for b in range(B)])] == 0)
    np.testing.assert_allclose(metrics.categorical_crossentropy(t, t), -np.log(1 - 1e-7), rtol=1e-3, atol=2e-7)
    np.testing.assert_allclose(d.dice(t, p), cport.dice(t, p), rtol=1e-6, atol=1e-7)
