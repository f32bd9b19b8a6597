"""
CPU tests: the C/OpenMP oracle (oracle/c, used for full-size checks and as bench.py's CPU
baseline) against the golden vectors and the numpy oracle -- bit-exact for interpolation.
"""
import numpy as np
import pytest

from conftest import golden_names, load_golden
from oracle import cport, interp

F32 = np.float32


def _fill(g):
    f = float(g['fill'])
    return None if np.isnan(f) else f


@pytest.mark.parametrize('name', golden_names('st_'))
def test_c_warp_golden_bit_exact(name):
    g = load_golden(name)
    np.testing.assert_array_equal(cport.warp(g['vol'], g['flow'], str(g['method']), _fill(g)), g['out'])


@pytest.mark.parametrize('name', [n for n in golden_names('interpn_3d_c') if 'cnone' not in n] +
                         ['interpn_1d_linear', 'interpn_1d_nearest', 'interpn_4d_linear_fillnone', 'interpn_4d_nearest_fill0p0',
                          'interpn_5d_linear_fill2p5'])
def test_c_interpn_golden_bit_exact(name):
    g = load_golden(name)
    vol = g['vol'] if g['vol'].ndim == g['loc'].shape[-1] + 1 else g['vol'][..., None]
    out = cport.interpn(vol, g['loc'], str(g['method']), _fill(g))
    ref = g['out'] if g['out'].ndim == out.ndim else g['out'][..., None]
    np.testing.assert_array_equal(out, ref)


def test_c_matches_numpy_on_a_larger_random_case():
    rng = np.random.default_rng(2)
    vol = rng.standard_normal((2, 24, 28, 40, 2)).astype(F32)
    flow = rng.uniform(-5, 5, (2, 24, 28, 40, 3)).astype(F32)
    for method, fill in (('linear', None), ('linear', 1.5), ('nearest', 0.0)):
        np.testing.assert_array_equal(cport.warp(vol, flow, method, fill),
                                      interp.spatial_transformer(vol, flow, method, 'ij', fill))


def test_c_dice_cce_lc3d():
    g = load_golden('dice_soft_default')
    np.testing.assert_allclose(cport.dice(g['y_true'], g['y_pred']), g['dice'], rtol=1e-6)
    g = load_golden('cce_label_weights')
    np.testing.assert_allclose(cport.cce(g['y_true'], g['y_pred'], g['label_weights']), g['loss'], rtol=1e-6)
    g = load_golden('lc3d_k3_s1_16to16')
    out = cport.lc3d(g['x'], g['kernel'], g['bias'].reshape(-1, 16), (3, 3, 3))
    np.testing.assert_allclose(out, g['out'], rtol=1e-5, atol=2e-5)
    g = load_golden('lc3d_k321_s212_cl')
    out = cport.lc3d(g['x'], g['kernel'], g['bias'].reshape(-1, 5), tuple(g['kernel_size']), tuple(g['strides']))
    np.testing.assert_allclose(out, g['out'], rtol=1e-5, atol=2e-5)


def test_c_mi_and_blur_match_golden_and_numpy_oracle():
    """the C/OpenMP restatements used as bench.py's cpu_baseline for --op mi / --op blur."""
    from oracle import conv as oconv, mi as omi
    g = load_golden('mi_channelwise_c3')
    np.testing.assert_allclose(cport.mi_channelwise(g['x'], g['y']), g['mi'], rtol=1e-5, atol=2e-6)
    g = load_golden('mi_volumes_nb16')
    np.testing.assert_allclose(cport.mi_channelwise(g['x'], g['y'], nb_bins=16).reshape(-1), g['mi'], rtol=1e-5, atol=2e-6)
    g = load_golden('mi_volumes_clip_alpha')
    out = cport.mi_channelwise(g['x'], g['y'], nb_bins=12, alpha=float(g['alpha']), min_clip=float(g['min_clip']),
                               max_clip=float(g['max_clip']))
    np.testing.assert_allclose(out.reshape(-1), g['mi'], rtol=1e-5, atol=2e-6)
    centers = np.linspace(0, 1, 9).astype(np.float32)
    rng = np.random.default_rng(2)
    x = rng.uniform(0, 1, (2, 300, 2)).astype(np.float32)
    y = rng.uniform(0, 1, (2, 300, 2)).astype(np.float32)
    np.testing.assert_allclose(cport.mi_channelwise(x, y, bin_centers=centers, alpha=30.0),
                               omi.MutualInformation(bin_centers=centers, soft_bin_alpha=30.0).channelwise(x, y),
                               rtol=1e-5, atol=2e-6)
    for name in ('blur3d_s1', 'blur3d_aniso', 'blur3d_s3', 'blur2d_s2', 'blur1d_s1p2'):
        g = load_golden(name)
        np.testing.assert_allclose(cport.gaussian_blur(g['x'], g['sigma'].tolist()), g['out'], rtol=1e-6, atol=1e-6)
    v = rng.standard_normal((1, 20, 21, 22, 1)).astype(np.float32)
    np.testing.assert_array_equal(cport.gaussian_blur(v, 1.3), oconv.gaussian_blur(v, 1.3))
