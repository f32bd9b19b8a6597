"""
LocallyConnected3D implementations 2 and 3 (reference layers.py:986-1028, 1260-1343) are storage variants of the
implementation-1 map.  The layer keeps the reference's parameter shapes / orderings and re-indexes them for the
streaming kernel (neurite_b200.layers.lc3d_kernel_from_impl / lc3d_kernel_to_impl).  Pinned here, on the CPU, by the
reference's own index generator: tests/golden/lc3d_impl_idx_* hold sorted(conv_kernel_idxs(...)) executed from the
reference source (tools/gen_golden.py gen_lc3d_impl).
"""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import lc3d as olc3d

F32 = np.float32


@pytest.mark.parametrize('name', ['lc3d_impl_idx_cl', 'lc3d_impl_idx_cl_s2', 'lc3d_impl_idx_cf'])
def test_impl2_impl3_weight_layouts_match_the_reference_index_order(name):
    from neurite_b200 import layers
    g = load_golden(name)
    idxs = g['idxs']
    I = tuple(int(v) for v in g['input_shape'])
    Cin, Cout = int(g['filters_in']), int(g['filters_out'])
    ks, st = tuple(int(v) for v in g['kernel_size']), tuple(int(v) for v in g['strides'])
    fmt = str(g['data_format'])
    O = tuple((I[d] - ks[d]) // st[d] + 1 for d in range(3))
    P, F = int(np.prod(O)), int(np.prod(ks)) * Cin
    assert idxs.shape == (P * F * Cout, 2)
    rng = np.random.default_rng(1)
    k1 = rng.standard_normal((P, F, Cout)).astype(F32)
    x = rng.standard_normal((2,) + (I + (Cin,) if fmt == 'channels_last' else (Cin,) + I)).astype(F32)
    y1 = olc3d.locally_connected_3d(x, k1, None, ks, st, data_format=fmt)
    in_size, out_size = int(np.prod(I)) * Cin, P * Cout
    xf = x.reshape(2, -1).astype(np.float64)

    # implementation 3: the weight vector IS the value list of the sparse (out, in) matrix in sorted index order
    vec = layers.lc3d_kernel_to_impl(torch.from_numpy(k1), 3, I, Cin, ks, st, fmt)
    assert tuple(vec.shape) == (len(idxs),)
    M = np.zeros((out_size, in_size))
    M[idxs[:, 0], idxs[:, 1]] = vec.numpy()
    y3 = (M @ xf.T).T.reshape(y1.shape)
    np.testing.assert_allclose(y3, y1, rtol=1e-5, atol=1e-5)
    back = layers.lc3d_kernel_from_impl(vec, 3, I, Cin, Cout, ks, st, fmt)
    assert torch.equal(back, torch.from_numpy(k1))

    # implementation 2: dense (input..., output...) weight; its support is exactly the reference's connectivity
    dense = layers.lc3d_kernel_to_impl(torch.from_numpy(k1), 2, I, Cin, ks, st, fmt)
    d2 = dense.numpy().reshape(in_size, out_size)
    support = np.zeros((in_size, out_size), dtype=bool)
    support[idxs[:, 1], idxs[:, 0]] = True
    assert np.array_equal(d2 != 0, support)
    y2 = (xf @ d2.astype(np.float64)).reshape(y1.shape)
    np.testing.assert_allclose(y2, y1, rtol=1e-5, atol=1e-5)
    noisy = dense + (~torch.from_numpy(support.reshape(dense.shape))) * 7.0      # entries outside the mask are ignored
    assert torch.equal(layers.lc3d_kernel_from_impl(noisy, 2, I, Cin, Cout, ks, st, fmt), torch.from_numpy(k1))


def test_layer_keeps_the_reference_parameter_shapes():
    from neurite_b200 import layers
    lay = layers.LocallyConnected3D(3, (2, 3, 2), implementation=3)
    lay.build((None, 4, 5, 3, 2))
    assert tuple(lay.kernel.shape) == (3 * 3 * 2 * (2 * 3 * 2 * 2) * 3,) and lay.get_config()['implementation'] == 3
    lay = layers.LocallyConnected3D(3, (2, 3, 2), implementation=2, data_format='channels_first')
    lay.build((None, 2, 4, 5, 3))
    assert tuple(lay.kernel.shape) == (2, 4, 5, 3, 3, 3, 3, 2)
    with pytest.raises(NotImplementedError):
        layers.LocallyConnected3D(3, 3, implementation=2, padding='same')
    with pytest.raises(ValueError, match='only "valid" is supported if implementation is 1'):
        layers.LocallyConnected3D(3, 3, implementation=1, padding='same')
