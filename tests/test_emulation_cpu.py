"""
CPU emulation of device-side bookkeeping that is easy to get wrong and expensive to debug on
the GPU: the plane ring, z segments, halo tiles and kernel centring of `blur3d_fused_kernel`
(neurite_b200/csrc/nrt_conv.cu).  The loops below mirror the kernel statement for statement (one
python iteration per CTA / plane); the result must equal the oracle's three-pass blur.
"""
import numpy as np
import pytest

from oracle import conv

FTX, FTY = 64, 16          # kFTX, kFTY


def emulate_blur3d_fused(x, kz, ky, kx, K, sm_count=148, zsplit_force=None):
    B, Z, Y, X = x.shape
    R, AY, AX = K // 2, FTY + K - 1, FTX + K - 1
    out = np.full_like(x, np.nan)
    cols = ((X + FTX - 1) // FTX) * ((Y + FTY - 1) // FTY) * B
    zsplit = -(-sm_count * 5 // cols)                                 # nrt_blur3d_f32: segment choice
    zsplit = max(min(zsplit, Z // (4 * K) if Z // (4 * K) > 0 else 1), 1)
    if zsplit_force:
        zsplit = zsplit_force
    zlen = -(-Z // zsplit)
    zsplit = -(-Z // zlen)
    for bz in range(B * zsplit):                                      # blockIdx.z
        b = bz // zsplit
        zb = (bz - b * zsplit) * zlen
        ze = min(zb + zlen, Z)
        for by in range((Y + FTY - 1) // FTY):                        # blockIdx.y
            for bx in range((X + FTX - 1) // FTX):                    # blockIdx.x
                x0, y0 = bx * FTX, by * FTY
                ring = np.full((K, FTY * FTX), np.nan)
                slot = 0
                for zi in range(zb - R, ze + R):
                    if 0 <= zi < Z:
                        A = np.zeros((AY, AX))
                        ys = slice(max(y0 - R, 0), min(y0 - R + AY, Y))
                        xs = slice(max(x0 - R, 0), min(x0 - R + AX, X))
                        A[ys.start - (y0 - R):ys.stop - (y0 - R), xs.start - (x0 - R):xs.stop - (x0 - R)] = x[b, zi, ys, xs]
                        Bx = sum(kx[j] * A[:, j:j + FTX] for j in range(K))
                        ring[slot] = sum(ky[j] * Bx[j:j + FTY, :] for j in range(K)).ravel()
                    else:
                        ring[slot] = 0.0
                    zo = zi - R
                    if zb <= zo < ze:
                        acc = np.zeros(FTY * FTX)
                        sl = 0 if slot + 1 == K else slot + 1
                        for j in range(K):
                            acc += kz[j] * ring[sl]
                            sl = 0 if sl + 1 == K else sl + 1
                        acc = acc.reshape(FTY, FTX)
                        ny, nx = min(FTY, Y - y0), min(FTX, X - x0)
                        out[b, zo, y0:y0 + ny, x0:x0 + nx] = acc[:ny, :nx]
                    slot = 0 if slot + 1 == K else slot + 1
    return out


@pytest.mark.parametrize('shape,sigma,zsplit', [((1, 9, 17, 70), [0.5, 2.3, 1.0], None), ((2, 30, 20, 65), 1.0, None),
                                                ((1, 64, 5, 3), [1.5, 0.0, 0.7], 2), ((1, 5, 4, 3), 2.0, None),
                                                ((1, 61, 16, 64), 0.7, 3)])
def test_fused_blur_bookkeeping_equals_three_passes(shape, sigma, zsplit):
    rng = np.random.default_rng(0)
    x = rng.standard_normal(shape)
    sg = sigma if isinstance(sigma, list) else [sigma] * 3
    ks = conv.gaussian_kernel(sg, separate=True)
    K = max(3, max(len(k) for k in ks))
    assert K % 2 == 1 and K <= 15
    pk = [np.pad(k.astype(np.float64), ((K - len(k)) // 2,) * 2) for k in ks]     # utils.pad_kernels_centered
    out = emulate_blur3d_fused(x, *pk, K, zsplit_force=zsplit)
    ref = conv.gaussian_blur(x[..., None].astype(np.float32), sg)[..., 0]
    assert not np.isnan(out).any()
    np.testing.assert_allclose(out, ref, rtol=0, atol=1e-6)
