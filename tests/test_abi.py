"""
CPU tests of the drop-in boundary: the C-ABI library loads here (no GPU), exports every
symbol include/neurite_b200.h declares, the ctypes binding covers the header one to one,
argument validation returns status codes (never throws, never touches the device), and the
python wrappers refuse CPU tensors instead of falling back.
"""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT

HEADER = os.path.join(ROOT, 'include', 'neurite_b200.h')


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(nrt_[a-z0-9_]+)\s*\(', src)))


def test_library_loads_and_exports_every_declared_symbol():
    from neurite_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 14
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), 'libneurite_b200.so does not export %s' % n
    assert sorted(_lib.SIGNATURES) == names, 'ctypes binding and header diverge'
    assert _lib.lib.nrt_version() == 1
    assert _lib.lib.nrt_status_string(0) == b'ok'


def test_argument_validation_returns_status_codes():
    from neurite_b200 import _lib
    lib = _lib.lib
    null = ctypes.c_void_p(0)
    shape = _lib.i32_array([4, 4, 4])
    assert lib.nrt_interpn_f32(null, shape, 3, 1, null, 0, 0, 0, 0.0, null, null) == -1      # null pointers
    assert b'null' in lib.nrt_last_error_string()
    one = ctypes.c_void_p(16)
    assert lib.nrt_interpn_f32(one, shape, 4, 1, one, 0, 0, 0, 0.0, one, null) == -1         # D = 4
    assert lib.nrt_interpn_f32(one, shape, 3, 1, one, 0, 7, 0, 0.0, one, null) == -1         # bad method
    assert b'linear or nearest' in lib.nrt_last_error_string()
    assert lib.nrt_warp_f32(one, one, one, 1, shape, 3, 1, 0, 0, 0.0, 0, 9, 0, 4, 0, null, null) == -1   # src planes > volume
    assert lib.nrt_dice_workspace_bytes(4, 16) == 4 * 2048 * 3 * 16 * 4
    assert lib.nrt_dice_sums_f32(one, one, 1, 10, 4, 5, 9, 0, 0, one, null, one, 1 << 30, null) == -1   # voxel range
    k = _lib.i32_array([3, 3, 3])
    assert lib.nrt_lc3d_fwd_f32(one, one, null, one, 1, _lib.i32_array([2, 8, 8]), 1, 1, k, k, 0, 0, 0, 1, null) == -1
    with pytest.raises(_lib.NeuriteB200Error, match='bad argument'):
        _lib.check(-1)


def test_no_cpu_fallback():
    import neurite_b200 as ne
    vol = torch.zeros(4, 4, 4)
    loc = torch.zeros(2, 3)
    with pytest.raises(ne._lib.NeuriteB200Error, match='no CPU path'):
        ne.utils.interpn(vol, loc)
    with pytest.raises(ne._lib.NeuriteB200Error, match='no CPU path'):
        ne.losses.Dice().loss(torch.zeros(1, 4, 4, 2), torch.zeros(1, 4, 4, 2))
    with pytest.raises(ne._lib.NeuriteB200Error, match='no CPU path'):
        ne.layers.SpatialTransformer()([torch.zeros(1, 4, 4, 4, 1), torch.zeros(1, 4, 4, 4, 3)])


def test_reference_exceptions_and_config_round_trip():
    import neurite_b200 as ne
    with pytest.raises(Exception, match='does not match volume dimension'):
        ne.utils.interpn(torch.zeros(3, 3, 3, 3, 3), torch.zeros(2, 2))
    with pytest.raises(AssertionError, match='method should be linear or nearest'):
        ne.utils.interpn(torch.zeros(3, 3), torch.zeros(2, 2), 'cubic')
    with pytest.raises(AssertionError):
        ne.utils.resize(torch.zeros(3), [2, 2])
    with pytest.raises(AssertionError):
        ne.metrics.Dice(input_type='one_hot')
    with pytest.raises(AssertionError, match='need nb_labels'):
        ne.metrics.Dice(dice_type='hard', input_type='max_label')
    with pytest.raises(ValueError, match='only "valid" is supported'):
        ne.layers.LocallyConnected3D(4, 3, padding='same')
    with pytest.raises(ValueError, match='Unrecognized implementation mode'):
        ne.layers.LocallyConnected3D(4, 3, implementation=4)
    with pytest.raises(ValueError, match='Label weights must be of len 3, but got 2'):
        ne.losses.CategoricalCrossentropy(label_weights=[1., 2.]).loss(torch.zeros(1, 2, 3), torch.zeros(1, 2, 3))
    r = ne.layers.Resize(2, interp_method='nearest')
    cfg = r.get_config()
    assert cfg['zoom_factor'] == 2 and cfg['interp_method'] == 'nearest'
    r.build((None, 8, 9, 10, 2))
    assert r.compute_output_shape((None, 8, 9, 10, 2)) == (None, 16, 18, 20, 2)
    with pytest.raises(AssertionError, match='zoom factor length'):
        ne.layers.Resize([2, 2]).build((None, 8, 9, 10, 2))
    lc = ne.layers.LocallyConnected3D(64, (3, 3, 3))
    lc.build((None, 32, 32, 32, 3))
    # docstring known answer, layers.py:825-835
    assert lc.compute_output_shape((None, 32, 32, 32, 3)) == (None, 30, 30, 30, 64)
    assert tuple(lc.kernel.shape) == (30 ** 3, 27 * 3, 64) and tuple(lc.bias.shape) == (30, 30, 30, 64)
    assert sorted(dict(lc.named_parameters())) == ['bias', 'kernel']
    cfg = lc.get_config()
    assert cfg['filters'] == 64 and cfg['kernel_size'] == (3, 3, 3) and cfg['implementation'] == 1
    assert ne.utils.flatten_axes(torch.zeros(3, 4, 5, 6), [1, 2]).shape == (3, 20, 6)   # utils.py:1200-1201
