"""
CPU tests: the oracle (oracle/*.py, numpy restatement) against
  * tests/golden/*.npz -- outputs of the reference's own source executed on tools/tfshim
    (tools/gen_golden.py), bit-exact for the interpolation family, 1e-6 for the sums;
  * the docstring known answers the reference carries;
  * independent implementations (scipy map_coordinates, torch grid_sample, F.conv3d).
"""
import numpy as np
import pytest

from conftest import golden_names, load_golden
from oracle import interp, lc3d, metrics

F32 = np.float32


def _fill(g):
    f = float(g['fill'])
    return None if np.isnan(f) else f


# ------------------------------------------------------------------ interpn vs golden
@pytest.mark.parametrize('name', golden_names('interpn_'))
def test_interpn_golden_bit_exact(name):
    g = load_golden(name)
    fill = _fill(g) if 'fill' in g.files else None
    loc = g['loc']
    if '_list_' in name:
        loc = [loc[..., d] for d in range(loc.shape[-1])]
    out = interp.interpn(g['vol'], loc, str(g['method']), fill)
    assert out.shape == g['out'].shape and out.dtype == g['out'].dtype
    np.testing.assert_array_equal(out, g['out'])


@pytest.mark.parametrize('name', golden_names('resize_'))
def test_resize_golden_bit_exact(name):
    g = load_golden(name)
    z = g['zoom']
    z = float(z) if z.ndim == 0 else [float(v) for v in z]
    if isinstance(z, float) and z.is_integer():
        z = int(z)
    if name.startswith('resize_layer'):
        out = interp.resize_layer(g['x'], z, str(g['method']))
    else:
        out = interp.resize(g['vol'], z, str(g['method']))
    np.testing.assert_array_equal(out, g['out'])


@pytest.mark.parametrize('name', golden_names('st_'))
def test_spatial_transformer_golden_bit_exact(name):
    g = load_golden(name)
    out = interp.spatial_transformer(g['vol'], g['flow'], str(g['method']), 'ij', _fill(g))
    np.testing.assert_array_equal(out, g['out'])


# ------------------------------------------------------------------ hand known answers
def test_interpn_known_answers():
    assert interp.interpn(np.array([[0, 1], [2, 3]], F32), np.array([[0.5, 0.5]], F32))[0] == F32(1.5)
    r = interp.interpn(np.arange(5, dtype=F32), np.array([[.5], [1.5], [2.5], [3.5]], F32), 'nearest')
    np.testing.assert_array_equal(r, [0, 2, 2, 4])          # half-to-even, utils.py:196
    # meshgrid docstring example, utils.py:413-424
    X, Y = interp.meshgrid(np.array([1, 2, 3]), np.array([4, 5, 6]))
    np.testing.assert_array_equal(X, [[1, 2, 3]] * 3)
    np.testing.assert_array_equal(Y, [[4] * 3, [5] * 3, [6] * 3])


def test_identity_warp_is_bit_exact_and_edges_clamp():
    rng = np.random.default_rng(3)
    vol = rng.standard_normal((1, 5, 6, 7, 2)).astype(F32)
    out = interp.spatial_transformer(vol, np.zeros((1, 5, 6, 7, 3), F32))
    np.testing.assert_array_equal(out, vol)
    far = np.full((1, 5, 6, 7, 3), 100, F32)
    out = interp.spatial_transformer(vol, far)
    np.testing.assert_array_equal(out, np.broadcast_to(vol[:, -1:, -1:, -1:], vol.shape))
    out = interp.spatial_transformer(vol, far, fill_value=-1)
    assert np.all(out[:, :-1] == -1) or np.all(out == -1)


def test_errors_match_reference():
    with pytest.raises(Exception, match='does not match volume dimension'):
        interp.interpn(np.zeros((3, 3, 3, 3, 3), F32), np.zeros((2, 2), F32))
    with pytest.raises(AssertionError, match='method should be linear or nearest'):
        interp.interpn(np.zeros((3, 3), F32), np.zeros((2, 2), F32), 'cubic')
    with pytest.raises(AssertionError):
        interp.resize(np.zeros((3,), F32), [2, 2])


# ------------------------------------------------------------------ independent cross-checks
def test_interpn_vs_scipy_and_grid_sample():
    from scipy.ndimage import map_coordinates
    import torch
    g = load_golden('interpn_cfg1_linear_32')
    vol, loc = g['vol'], g['loc']
    ours = interp.interpn(vol, loc)
    sp = map_coordinates(vol.astype(np.float64), np.moveaxis(loc, -1, 0).astype(np.float64), order=1, mode='nearest')
    assert np.max(np.abs(ours - sp)) < 2e-6
    # torch: grid is xyz-reversed, normalised to [-1,1], align_corners=True, border padding
    S = np.array(vol.shape, dtype=np.float64)
    grid = (2 * loc.astype(np.float64) / (S - 1) - 1)[..., ::-1].copy()
    ts = torch.nn.functional.grid_sample(torch.from_numpy(vol)[None, None].double(), torch.from_numpy(grid)[None],
                                         mode='bilinear', padding_mode='border', align_corners=True)[0, 0].numpy()
    assert np.max(np.abs(ours - ts)) < 5e-6


# ------------------------------------------------------------------ Dice / CCE
@pytest.mark.parametrize('name', ['dice_soft_default', 'dice_soft_laplace_weights', 'dice_soft_normalize',
                                  'dice_hard_prob', 'dice_hard_max_label', 'dice_soft_disjoint_L5'])
def test_dice_golden(name):
    g = load_golden(name)
    kw = {}
    if name == 'dice_soft_laplace_weights':
        kw = dict(weights=g['weights'], laplace_smoothing=0.1)
    elif name == 'dice_soft_normalize':
        kw = dict(normalize=True)
    elif name == 'dice_hard_prob':
        kw = dict(dice_type='hard', input_type='prob')
    elif name == 'dice_hard_max_label':
        kw = dict(dice_type='hard', input_type='max_label', nb_labels=int(g['nb_labels']))
    with np.errstate(all='ignore'):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            d = metrics.Dice(**kw)
            np.testing.assert_allclose(d.dice(g['y_true'], g['y_pred']), g['dice'], rtol=1e-6, atol=1e-7)
            for k, fn in (('loss', d.loss), ('mean_dice', d.mean_dice), ('mean_loss', d.mean_loss)):
                if k in g.files:
                    np.testing.assert_allclose(fn(g['y_true'], g['y_pred']), g[k], rtol=1e-6, atol=1e-7)


def test_dice_range_error_and_properties():
    g = load_golden('dice_range_error')
    assert 'value outside range' in str(g['raised'])
    with pytest.raises(metrics.RangeError, match='value outside range'):
        metrics.Dice().dice(g['y_true'], g['y_pred'])
    t = np.eye(4, dtype=F32)[np.random.default_rng(0).integers(0, 4, (2, 5, 5))]
    np.testing.assert_array_equal(metrics.Dice().dice(t, t), np.ones((2, 4), F32))
    with pytest.raises(AssertionError):
        metrics.Dice(input_type='one_hot')                       # documented but rejected, metrics.py:406
    with pytest.raises(AssertionError):
        metrics.Dice(dice_type='hard', input_type='max_label')


@pytest.mark.parametrize('name', ['cce_label_weights', 'cce_plain'])
def test_cce_golden(name):
    g = load_golden(name)
    lw = g['label_weights'] if 'label_weights' in g.files else None
    out = metrics.categorical_crossentropy(g['y_true'], g['y_pred'], label_weights=lw)
    np.testing.assert_allclose(out, g['loss'], rtol=1e-6)


def test_cce_known_answer_and_error():
    g = load_golden('cce_bad_weights')
    assert 'Label weights must be of len 16, but got 5' in str(g['raised'])
    t = np.eye(3, dtype=F32)[[0, 1, 2, 1]][None]
    np.testing.assert_allclose(metrics.categorical_crossentropy(t, t), -np.log(1 - 1e-7), rtol=1e-3, atol=2e-7)
    with pytest.raises(ValueError, match='Label weights must be of len 3'):
        metrics.categorical_crossentropy(t, t, label_weights=np.ones(4))


# ------------------------------------------------------------------ LocallyConnected3D
@pytest.mark.parametrize('name', [n for n in golden_names('lc3d_') if 'impl_idx' not in n])
def test_lc3d_golden(name):
    g = load_golden(name)
    bias = g['bias'] if g['bias'].size else None
    for literal in (True, False):
        out = lc3d.locally_connected_3d(g['x'], g['kernel'], bias, tuple(g['kernel_size']), tuple(g['strides']),
                                        'valid', str(g['data_format']), None, literal=literal)
        assert out.shape == g['out'].shape
        np.testing.assert_allclose(out, g['out'], rtol=1e-5, atol=2e-5)


def test_lc3d_shapes_and_shared_weights_equal_conv3d():
    import torch
    # docstring: (32,32,32,3) -> k3 -> (30,30,30,64) -> k3 -> (28,28,28,32)   layers.py:825-835
    assert lc3d.output_shape((32, 32, 32), (3, 3, 3), (1, 1, 1)) == (30, 30, 30)
    assert lc3d.output_shape((30, 30, 30), (3, 3, 3), (1, 1, 1)) == (28, 28, 28)
    assert 30 ** 3 * (27 * 3 * 64) + 30 ** 3 * 64 == 30 ** 3 * 64 * (81 + 1)
    rng = np.random.default_rng(1)
    x = rng.standard_normal((2, 6, 6, 6, 3)).astype(F32)
    w = rng.standard_normal((3, 3, 3, 3, 4)).astype(F32)                 # [k0,k1,k2,Cin,Cout]
    P = 4 ** 3
    kernel = np.broadcast_to(w.reshape(1, 81, 4), (P, 81, 4)).copy()
    ours = lc3d.locally_connected_3d(x, kernel, None, (3, 3, 3))
    ref = torch.nn.functional.conv3d(torch.from_numpy(x).permute(0, 4, 1, 2, 3),
                                     torch.from_numpy(w).permute(4, 3, 0, 1, 2)).permute(0, 2, 3, 4, 1).numpy()
    np.testing.assert_allclose(ours, ref, rtol=1e-4, atol=1e-4)
    with pytest.raises(ValueError, match='only "valid" is supported'):
        lc3d.locally_connected_3d(x, kernel, None, (3, 3, 3), padding='same')
