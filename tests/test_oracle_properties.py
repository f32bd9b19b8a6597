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


# ------------------------------------------------------------------ MutualInformation / convolution oracles
from oracle import conv as oconv, mi as omi   # noqa: E402


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 2**31 - 1), st.integers(3, 20), st.integers(1, 60))
def test_mi_symmetry_and_voxel_permutation_invariance(seed, nb, nvox):
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, 1, (2, nvox, 1)).astype(np.float32)
    y = rng.uniform(0, 1, (2, nvox, 1)).astype(np.float32)
    m = omi.MutualInformation(nb_bins=nb)
    a = m.volumes(x, y)
    np.testing.assert_allclose(a, m.volumes(y, x), rtol=1e-5, atol=2e-6)          # MI(x, y) = MI(y, x)
    perm = rng.permutation(nvox)
    np.testing.assert_allclose(a, m.volumes(x[:, perm], y[:, perm]), rtol=1e-5, atol=2e-6)
    assert a.shape == (2,) and np.all(np.isfinite(a))


@settings(max_examples=15, deadline=None)
@given(st.integers(0, 2**31 - 1), st.integers(2, 12))
def test_mi_maps_of_independent_product_is_zero_and_identical_onehot_is_entropy(seed, L):
    rng = np.random.default_rng(seed)
    # x constant over voxels -> the joint factorises exactly -> MI = 0 (up to the eps terms)
    p = rng.dirichlet(np.ones(L), size=(1, 50)).astype(np.float32)
    q = np.broadcast_to(rng.dirichlet(np.ones(L)).astype(np.float32), (1, 50, L)).copy()
    assert abs(float(omi.MutualInformation().maps(q, p)[0])) < 1e-4
    # identical one-hot maps: MI = entropy of the label distribution
    lab = rng.integers(0, L, 400)
    oh = np.eye(L, dtype=np.float32)[lab][None]
    cnt = np.bincount(lab, minlength=L) / 400.0
    ent = -np.sum(cnt[cnt > 0] * np.log(cnt[cnt > 0]))
    np.testing.assert_allclose(omi.MutualInformation().maps(oh, oh)[0], ent, rtol=1e-4, atol=1e-5)


@settings(max_examples=20, deadline=None)
@given(st.integers(0, 2**31 - 1), st.sampled_from([0.5, 1.0, 1.7, 2.4]), st.sampled_from([0.0, 0.8, 1.3]))
def test_blur_oracle_is_linear_separable_and_mass_preserving(seed, s0, s1):
    rng = np.random.default_rng(seed)
    shape = (1, 14, 15, 2)
    a = rng.standard_normal(shape).astype(np.float32)
    b = rng.standard_normal(shape).astype(np.float32)
    sig = [s0, s1]
    lin = oconv.gaussian_blur(a + 2 * b, sig)
    np.testing.assert_allclose(lin, oconv.gaussian_blur(a, sig) + 2 * oconv.gaussian_blur(b, sig), rtol=0, atol=2e-5)
    # the passes commute (each is a zero-padded cross-correlation along its own axis)
    ks = oconv.gaussian_kernel(sig, separate=True)
    ks = ks if isinstance(ks, list) else [ks]
    f01 = oconv.conv1d_axis(oconv.conv1d_axis(a, ks[0], 1), ks[1], 2)
    f10 = oconv.conv1d_axis(oconv.conv1d_axis(a, ks[1], 2), ks[0], 1)
    np.testing.assert_allclose(f01, f10, rtol=0, atol=1e-5)
    np.testing.assert_allclose(f01, oconv.gaussian_blur(a, sig), rtol=0, atol=1e-6)
    # an impulse far from the border keeps its mass (the kernel sums to 1)
    imp = np.zeros((1, 31, 33, 1), np.float32)
    imp[0, 15, 16, 0] = 3.0
    assert abs(float(oconv.gaussian_blur(imp, sig).sum()) - 3.0) < 1e-5


@settings(max_examples=20, deadline=None)
@given(st.integers(0, 2**31 - 1), st.integers(1, 9), st.integers(1, 3), st.integers(1, 3))
def test_conv1d_axis_matches_scipy_correlate_and_tf_shape_rules(seed, K, stride, dil):
    from scipy import ndimage
    if stride > 1 and dil > 1:
        dil = 1
    rng = np.random.default_rng(seed)
    n = 23
    x = rng.standard_normal((2, n, 3)).astype(np.float32)
    k = rng.standard_normal(K).astype(np.float32)
    same = oconv.conv1d_axis(x, k, 1, 'SAME', stride, dil)
    valid = oconv.conv1d_axis(x, k, 1, 'VALID', stride, dil)
    assert same.shape[1] == -(-n // stride)                                      # ceil(n / stride)
    assert valid.shape[1] == max(-(-(n - (K - 1) * dil) // stride), 0)
    if K % 2 == 1 and stride == 1:
        kd = np.zeros((K - 1) * dil + 1, np.float64)
        kd[::dil] = k
        ref = ndimage.correlate1d(x.astype(np.float64), kd, axis=1, mode='constant', cval=0.0)
        np.testing.assert_allclose(same, ref, rtol=1e-5, atol=1e-5)
        r = (len(kd) - 1) // 2
        if valid.shape[1] > 0:
            np.testing.assert_allclose(valid, ref[:, r:n - r], rtol=1e-5, atol=1e-5)
