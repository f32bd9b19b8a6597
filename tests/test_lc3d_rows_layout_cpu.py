"""
CPU model of lc3d_rows_kernel's lane algorithm (neurite_b200/csrc/nrt_lc3d.cu): which (row, chunk) a lane reads when,
which accumulator holds what, and the select-free fold.  The kernel itself is tested on the GPU against the oracle
(tests/test_gpu_parity.py::test_lc3d_row_kernel_vs_oracle); this file pins the three invariants the design rests on,
with plain numpy and no device:

  * the read order q ^ m, m = (lane >> 1) & 3, makes every quarter-warp of a 128-bit shared-memory load hit eight
    distinct 16-byte bank groups although the rows are 64 bytes apart,
  * batch slot b of a lane accumulating item b ^ (the lane's own item) lets every lane keep its LOWER slots at each
    batch step of the reduce-scatter, and register set d holds exactly the chunk that lane l ^ 2d ends up with,
  * every (item, chunk) output is written exactly once and equals the plain matrix product.
"""
import numpy as np
import pytest


def lane_maps(BB):
    lanes = np.arange(32)
    m = (lanes >> 1) & 3
    if BB == 8:
        pb = ((lanes >> 4) & 1) * 4 + ((lanes >> 3) & 1) * 2 + (lanes & 1)
    else:
        pb = ((lanes >> 4) & 1) * 2 + ((lanes >> 3) & 1)
    return m, pb


def fold_upper(v, n, mask):
    new = v.copy()
    for lane in range(32):
        new[lane, :n // 2] = v[lane, :n // 2] + v[lane ^ mask, n // 2:n]
    return new


def rows_kernel_model(x, w, BB):
    """x [BB, F], w [F, 16] -> out [BB, 16] computed the way one warp of lc3d_rows_kernel<BB, .> does"""
    F = w.shape[0]
    m, pb = lane_maps(BB)
    acc = np.zeros((32, BB, 4, 4))                       # lane, batch slot, register set, channel in chunk
    for c in range((F + 31) // 32):
        for lane in range(32):
            j = lane + 32 * c
            if j >= F:
                continue                                   # ragged last step
            for q in range(4):
                chunk = q ^ m[lane]
                for b in range(BB):
                    acc[lane, b, q] += x[b ^ pb[lane], j] * w[j, 4 * chunk:4 * chunk + 4]
    v = acc.reshape(32, BB * 16)
    v = fold_upper(v, BB * 16, 16)
    v = fold_upper(v, BB * 8, 8)
    if BB == 8:
        v = fold_upper(v, BB * 4, 1)
    u = np.zeros((32, 4))
    for lane in range(32):
        u[lane] = v[lane, 0:4] + v[lane ^ 2, 4:8] + v[lane ^ 4, 8:12] + v[lane ^ 6, 12:16]
    if BB == 4:
        u = u + u[np.arange(32) ^ 1]
    out = np.full((BB, 16), np.nan)
    writes = np.zeros((BB, 4), dtype=int)
    for lane in range(32):
        if BB == 8 or (lane & 1) == 0:
            out[pb[lane], 4 * m[lane]:4 * m[lane] + 4] = u[lane]
            writes[pb[lane], m[lane]] += 1
    return out, writes


@pytest.mark.parametrize('BB', [4, 8])
@pytest.mark.parametrize('F', [432, 48, 20])
def test_rows_kernel_model_equals_matrix_product(BB, F):
    rng = np.random.default_rng(BB * 1000 + F)
    x = rng.standard_normal((BB, F))
    w = rng.standard_normal((F, 16))
    out, writes = rows_kernel_model(x, w, BB)
    assert (writes == 1).all()                           # every (item, chunk) has exactly one owner
    np.testing.assert_allclose(out, x @ w, rtol=1e-12, atol=1e-12)


def test_chunk_order_is_bank_conflict_free():
    m, _ = lane_maps(4)
    for q in range(4):                                   # the q-th LDS.128 of a row step
        for quarter in range(4):                         # a 128-bit warp load is served a quarter-warp at a time
            lanes = np.arange(8) + 8 * quarter
            float4_index = lanes * 4 + (q ^ m[lanes])    # row = lane (+ 32 c: adds a multiple of 128 float4)
            assert len(set(float4_index % 8)) == 8       # eight distinct 16-byte bank groups
    # without the permutation the same load is 4-way conflicted
    assert len(set((np.arange(8) * 4) % 8)) == 2


def test_fold_partners_share_the_chunk_permutation():
    m, _ = lane_maps(8)
    for mask in (16, 8, 1):                              # the batch steps of the fold
        assert (m == m[np.arange(32) ^ mask]).all()
    for d in (1, 2, 3):                                  # register set d of lane l holds the own chunk of lane l ^ 2d
        assert ((m ^ d) == m[np.arange(32) ^ (2 * d)]).all()
