// nrt_mi_bwd.cu -- gradient of the mutual-information metric (SURVEY.md 8f-1 for the 8f-3 row).
//
// The reference gets this from TF autodiff through metrics.py:227-292 (maps) and
// utils.py:1099-1172 (soft_quantize); the chain, written out:
//   mi = f(H, sx, sy)                               (nrt_mi_finalize_f32)
//   H[i][j] = sum_v wx_i(v) wy_j(v),  sx[i] = sum_v wx_i(v),  sy[j] = sum_v wy_j(v)
//   wx_i(v) = exp(-alpha (clip(x_v) - cx_i)^2)      (quantised operand)   or   x[v][i]   (map operand)
//   cx_i    = linspace(min x, max x, nb)[i]         (when the centres are not given)
// 1. nrt_mi_finalize_bwd_f32:  G = dL/dH, gsx = dL/dsx, gsy = dL/dsy per item (fp64 inside).
// 2. nrt_mi_bwd_f32: one pass over the voxels,
//      T_i(v) = gsx_i + sum_j G_ij wy_j(v)          dL/dwx_i(v)
//      quantised: dL/dx_v = [min_clip <= x_v <= max_clip] * sum_i T_i wx_i (-2 alpha (clip(x_v) - cx_i))
//                 dL/dcx_i = - sum_v T_i wx_i (-2 alpha (clip(x_v) - cx_i))          (per-bin sums)
//      map:       dL/dx[v][i] = T_i(v)
//    and the same for y with G transposed.
// 3. nrt_mi_minmax_bwd_f32: the centres' gradient flows to the extrema,
//      dL/dmin = sum_i dL/dcx_i (1 - i/(nb-1)),  dL/dmax = sum_i dL/dcx_i i/(nb-1),
//    shared evenly by the elements equal to the minimum / maximum (TF's reduce_min/max gradient).
// CUDA cores only: 2 nb^2 FMAs per voxel with the G matrices broadcast from shared memory.
#include "nrt_common.cuh"

#include <math.h>

namespace nrt {
namespace {

constexpr int kBwdThreads = 256;
constexpr int kBwdMaxBlocks = 1184;

struct BwdOperand {
  const float* p;          // values
  float* g;                // gradient (same layout), may be null
  int64_t batch_stride, vox_stride;
  int nb, quant;
  const float* centers;
};
struct BwdArgs {
  BwdOperand x, y;
  int64_t nv;
  float neg_alpha, lo, hi;
  const float* gstats;     // [items][nbx*nby + nbx + nby]
  float* partial;          // [blocks][2 * (NBP + 2)]  (dcx, cnt_min, cnt_max, dcy, cnt_min, cnt_max) or null
};

// stats -> gstats, one block per item
__global__ void mi_finalize_bwd_kernel(const float* stats, const float* grad_mi, int nbx, int nby, float eps,
                                       float* gstats) {
  __shared__ double s_red[8];
  __shared__ double s_px[64], s_py[64], s_Px[64], s_Py[64];
  const int item = blockIdx.x;
  const int npair = nbx * nby, PS = npair + nbx + nby;
  const float* h = stats + (int64_t)item * PS;
  const float* sx = h + npair;
  const float* sy = sx + nbx;
  float* out = gstats + (int64_t)item * PS;
  const double g = (double)grad_mi[item];
  auto block_sum = [&](double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += s_red[w];
    return s;
  };
  double th = 0.0, tx = 0.0, ty = 0.0;
  for (int i = threadIdx.x; i < npair; i += blockDim.x) th += (double)h[i];
  for (int i = threadIdx.x; i < nbx; i += blockDim.x) tx += (double)sx[i];
  for (int i = threadIdx.x; i < nby; i += blockDim.x) ty += (double)sy[i];
  const double N = block_sum(th) + (double)eps, Nx = block_sum(tx) + (double)eps, Ny = block_sum(ty) + (double)eps;
  for (int i = threadIdx.x; i < 64; i += blockDim.x) {
    s_px[i] = i < nbx ? (double)sx[i] / Nx : 0.0;
    s_py[i] = i < nby ? (double)sy[i] / Ny : 0.0;
    s_Px[i] = 0.0;
    s_Py[i] = 0.0;
  }
  __syncthreads();
  // D = d mi / d pxy,  E = d mi / d q  with q = px py + eps, r = pxy / q + eps, mi = sum pxy log r
  double sd = 0.0;
  for (int idx = threadIdx.x; idx < npair; idx += blockDim.x) {
    const int i = idx / nby, j = idx - i * nby;
    const double pxy = (double)h[idx] / N;
    const double q = s_px[i] * s_py[j] + (double)eps;
    const double r = pxy / q + (double)eps;
    const double D = log(r) + pxy / (r * q);
    const double E = -pxy * pxy / (r * q * q);
    sd += D * pxy;
    atomicAdd(&s_Px[i], E * s_py[j]);
    atomicAdd(&s_Py[j], E * s_px[i]);
  }
  const double SD = block_sum(sd);
  double spx = 0.0, spy = 0.0;
  for (int i = threadIdx.x; i < nbx; i += blockDim.x) spx += s_Px[i] * s_px[i];
  for (int i = threadIdx.x; i < nby; i += blockDim.x) spy += s_Py[i] * s_py[i];
  const double SPx = block_sum(spx), SPy = block_sum(spy);
  for (int idx = threadIdx.x; idx < npair; idx += blockDim.x) {
    const int i = idx / nby, j = idx - i * nby;
    const double pxy = (double)h[idx] / N;
    const double q = s_px[i] * s_py[j] + (double)eps;
    const double r = pxy / q + (double)eps;
    const double D = log(r) + pxy / (r * q);
    out[idx] = (float)(g * (D - SD) / N);
  }
  for (int i = threadIdx.x; i < nbx; i += blockDim.x) out[npair + i] = (float)(g * (s_Px[i] - SPx) / Nx);
  for (int i = threadIdx.x; i < nby; i += blockDim.x) out[npair + nbx + i] = (float)(g * (s_Py[i] - SPy) / Ny);
}

// One operand's side of the voxel pass.  w* = this operand's weights (registers), o* = the other's.
//   T_i = gs[i] + sum_j M[i][j] wo[j]
template <int NBP, bool Q>
__device__ __forceinline__ void bwd_side(const BwdOperand& op, const float* vp, bool ok_grad, float xc, float two_neg_alpha,
                                         const float* M, const float* gs, const float* cen, const float (&wself)[NBP],
                                         const float (&wother)[NBP], float (&dc)[NBP], bool want_dc, float* gp) {
  float gsum = 0.f;
#pragma unroll
  for (int i = 0; i < NBP; ++i) {
    float T = gs[i];
    const float4* row = reinterpret_cast<const float4*>(M + i * NBP);
#pragma unroll
    for (int j4 = 0; j4 < NBP / 4; ++j4) {
      const float4 m = row[j4];                           // same address for the whole warp: one broadcast
      T = fmaf(m.x, wother[4 * j4 + 0], T);
      T = fmaf(m.y, wother[4 * j4 + 1], T);
      T = fmaf(m.z, wother[4 * j4 + 2], T);
      T = fmaf(m.w, wother[4 * j4 + 3], T);
    }
    if (Q) {
      const float A = wself[i] * (two_neg_alpha * (xc - cen[i])) * T;      // d/dx of T_i exp(-alpha (x - c_i)^2)
      gsum += A;
      if (want_dc) dc[i] -= A;
    } else if (i < op.nb) {
      gp[i] = T;
    }
  }
  if (Q) gp[0] = ok_grad ? gsum : 0.f;
}

// grid (nblk, C, B); smem: G [NBP][NBP], G^T [NBP][NBP], gsx, gsy, cx, cy [NBP each]
template <int NBP, bool QX, bool QY>
__global__ void __launch_bounds__(kBwdThreads) mi_bwd_voxel_kernel(const BwdArgs a, int want_dcx, int want_dcy) {
  extern __shared__ __align__(16) float smem[];
  float* Gs = smem;
  float* GTs = Gs + NBP * NBP;
  float* gsx = GTs + NBP * NBP;
  float* gsy = gsx + NBP;
  float* cxs = gsy + NBP;
  float* cys = cxs + NBP;
  __shared__ float s_part[8][2 * (NBP + 2)];
  const int item = blockIdx.z, chan = blockIdx.y;
  const int nbx = a.x.nb, nby = a.y.nb, npair = nbx * nby;
  const float* gst = a.gstats + (int64_t)(item * gridDim.y + chan) * (npair + nbx + nby);
  for (int e = threadIdx.x; e < NBP * NBP; e += kBwdThreads) {
    const int i = e / NBP, j = e - i * NBP;
    const float v = (i < nbx && j < nby) ? gst[i * nby + j] : 0.f;
    Gs[i * NBP + j] = v;
    GTs[j * NBP + i] = v;
  }
  for (int e = threadIdx.x; e < NBP; e += kBwdThreads) {
    gsx[e] = e < nbx ? gst[npair + e] : 0.f;
    gsy[e] = e < nby ? gst[npair + nbx + e] : 0.f;
    cxs[e] = (QX && e < nbx) ? a.x.centers[e] : 0.f;
    cys[e] = (QY && e < nby) ? a.y.centers[e] : 0.f;
  }
  __syncthreads();
  const float* xi = a.x.p + (int64_t)item * a.x.batch_stride + (QX ? chan : 0);
  const float* yi = a.y.p + (int64_t)item * a.y.batch_stride + (QY ? chan : 0);
  float* gxi = a.x.g ? a.x.g + (int64_t)item * a.x.batch_stride + (QX ? chan : 0) : nullptr;
  float* gyi = a.y.g ? a.y.g + (int64_t)item * a.y.batch_stride + (QY ? chan : 0) : nullptr;
  const float two_neg_alpha = 2.f * a.neg_alpha;
  const float xmin = QX ? cxs[0] : 0.f, xmax = QX ? cxs[nbx - 1] : 0.f;
  const float ymin = QY ? cys[0] : 0.f, ymax = QY ? cys[nby - 1] : 0.f;

  float dcx[NBP], dcy[NBP];
#pragma unroll
  for (int i = 0; i < NBP; ++i) dcx[i] = dcy[i] = 0.f;
  float cnt[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t v = (int64_t)blockIdx.x * kBwdThreads + threadIdx.x; v < a.nv; v += (int64_t)gridDim.x * kBwdThreads) {
    const float* xp = xi + v * a.x.vox_stride;
    const float* yp = yi + v * a.y.vox_stride;
    float wx[NBP], wy[NBP];
    float xc = 0.f, yc = 0.f;
    bool inx = true, iny = true;
    if (QX) {
      const float xr = xp[0];
      inx = (xr >= a.lo) && (xr <= a.hi);
      xc = fminf(fmaxf(xr, a.lo), a.hi);
      cnt[0] += (xr == xmin) ? 1.f : 0.f;
      cnt[1] += (xr == xmax) ? 1.f : 0.f;
#pragma unroll
      for (int i = 0; i < NBP; ++i) {
        const float d = xc - cxs[i];
        wx[i] = i < nbx ? __expf(a.neg_alpha * (d * d)) : 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < NBP; ++i) wx[i] = i < nbx ? xp[i] : 0.f;
    }
    if (QY) {
      const float yr = yp[0];
      iny = (yr >= a.lo) && (yr <= a.hi);
      yc = fminf(fmaxf(yr, a.lo), a.hi);
      cnt[2] += (yr == ymin) ? 1.f : 0.f;
      cnt[3] += (yr == ymax) ? 1.f : 0.f;
#pragma unroll
      for (int i = 0; i < NBP; ++i) {
        const float d = yc - cys[i];
        wy[i] = i < nby ? __expf(a.neg_alpha * (d * d)) : 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < NBP; ++i) wy[i] = i < nby ? yp[i] : 0.f;
    }
    if (gxi) bwd_side<NBP, QX>(a.x, xp, inx, xc, two_neg_alpha, Gs, gsx, cxs, wx, wy, dcx, want_dcx != 0, gxi + v * a.x.vox_stride);
    if (gyi) bwd_side<NBP, QY>(a.y, yp, iny, yc, two_neg_alpha, GTs, gsy, cys, wy, wx, dcy, want_dcy != 0, gyi + v * a.y.vox_stride);
  }
  if (!a.partial) return;
  // block sums of the per-bin centre gradients and of the tie counts
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NBP; ++i) {
    const float sxv = warp_sum(dcx[i]), syv = warp_sum(dcy[i]);
    if (lane == 0) { s_part[warp][i] = sxv; s_part[warp][NBP + 2 + i] = syv; }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float c = warp_sum(cnt[k]);
    if (lane == 0) s_part[warp][(k < 2 ? NBP + k : 2 * NBP + 2 + (k - 2))] = c;
  }
  __syncthreads();
  const int64_t blin = ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  for (int e = threadIdx.x; e < 2 * (NBP + 2); e += kBwdThreads) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += s_part[w][e];
    a.partial[blin * (2 * (NBP + 2)) + e] = s;
  }
}

// dcenters[o][0..nb) = sum over blocks of dc, [nb] = #min ties, [nb+1] = #max ties   (o = 0: x, 1: y)
__global__ void mi_bwd_combine_kernel(const float* partial, int64_t nblocks, int NBP, int nbx, int nby, float* dcenters,
                                      int stride_out) {
  const int e = threadIdx.x;                           // 0 .. 2 * (NBP + 2)
  if (e >= 2 * (NBP + 2)) return;
  double s = 0.0;
  for (int64_t b = 0; b < nblocks; ++b) s += (double)partial[b * (2 * (NBP + 2)) + e];
  const int o = e / (NBP + 2), k = e - o * (NBP + 2);
  const int nb = o == 0 ? nbx : nby;
  if (k < NBP) {
    if (k < nb) dcenters[o * stride_out + k] = (float)s;
  } else {
    dcenters[o * stride_out + nb + (k - NBP)] = (float)s;
  }
}

// grad_x[e] += dmin / #ties where x[e] == min, += dmax / #ties where x[e] == max
__global__ void mi_minmax_bwd_kernel(const float* x, int64_t n, const float* minmax, const float* dc, int nb, float* gx) {
  const float mn = minmax[0], mx = minmax[1];
  float dmin = 0.f, dmax = 0.f;
  for (int i = 0; i < nb; ++i) {
    const float f = nb > 1 ? (float)i / (float)(nb - 1) : 0.f;
    dmin += dc[i] * (1.f - f);
    dmax += dc[i] * f;
  }
  const float cmin = fmaxf(dc[nb], 1.f), cmax = fmaxf(dc[nb + 1], 1.f);
  dmin /= cmin;
  dmax /= cmax;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const float v = ld_stream_f(x + e);
    if (v == mn || v == mx) gx[e] += (v == mn ? dmin : 0.f) + (v == mx ? dmax : 0.f);
  }
}

template <int NBP>
void launch_bwd(const BwdArgs& a, dim3 grid, int want_dcx, int want_dcy, cudaStream_t st) {
  const size_t smem = (size_t)(2 * NBP * NBP + 4 * NBP) * sizeof(float);
  if (a.x.quant && a.y.quant) mi_bwd_voxel_kernel<NBP, true, true><<<grid, kBwdThreads, smem, st>>>(a, want_dcx, want_dcy);
  else if (a.x.quant) mi_bwd_voxel_kernel<NBP, true, false><<<grid, kBwdThreads, smem, st>>>(a, want_dcx, want_dcy);
  else if (a.y.quant) mi_bwd_voxel_kernel<NBP, false, true><<<grid, kBwdThreads, smem, st>>>(a, want_dcx, want_dcy);
  else mi_bwd_voxel_kernel<NBP, false, false><<<grid, kBwdThreads, smem, st>>>(a, want_dcx, want_dcy);
}

int bwd_blocks(int64_t nv, int items) {
  int64_t want = ((int64_t)sm_count() * 4 + items - 1) / items;
  int64_t cap = (nv + kBwdThreads - 1) / kBwdThreads;
  int64_t n = want < cap ? want : cap;
  if (n < 1) n = 1;
  if (n > kBwdMaxBlocks) n = kBwdMaxBlocks;
  return (int)n;
}

}  // namespace
}  // namespace nrt

using namespace nrt;

extern "C" {

int nrt_mi_finalize_bwd_f32(const float* stats, const float* grad_mi, int items, int nbx, int nby, float eps,
                            float* gstats, void* stream) {
  NRT_REQUIRE(stats && grad_mi && gstats, NRT_E_ARG, "null pointer");
  NRT_REQUIRE(items >= 1 && nbx >= 1 && nby >= 1 && nbx <= 64 && nby <= 64, NRT_E_ARG, "bad items/bins");
  mi_finalize_bwd_kernel<<<items, 128, 0, static_cast<cudaStream_t>(stream)>>>(stats, grad_mi, nbx, nby, eps, gstats);
  return check_launch("mi_finalize_bwd_kernel");
}

int64_t nrt_mi_bwd_workspace_bytes(int items) {
  if (items < 1) return 0;
  return (int64_t)items * kBwdMaxBlocks * 2 * (32 + 2) * (int64_t)sizeof(float);
}

int nrt_mi_bwd_f32(const float* x, int64_t x_batch_stride, int64_t x_vox_stride, int x_quant, int nbx,
                   const float* x_centers, const float* y, int64_t y_batch_stride, int64_t y_vox_stride,
                   int y_quant, int nby, const float* y_centers, int B, int C, int64_t nv, float alpha,
                   float min_clip, float max_clip, const float* gstats, float* grad_x, float* grad_y,
                   float* dcenters, void* workspace, int64_t workspace_bytes, void* stream) {
  NRT_REQUIRE(x && y && gstats, NRT_E_ARG, "null pointer");
  NRT_REQUIRE(grad_x || grad_y, NRT_E_ARG, "no gradient requested");
  NRT_REQUIRE(B >= 1 && C >= 1 && nv >= 0, NRT_E_ARG, "bad B/C/nv");
  NRT_REQUIRE(nbx >= 1 && nby >= 1 && nbx <= 32 && nby <= 32, NRT_E_SIZE, "the gradient supports up to 32 bins (got %d, %d)", nbx, nby);
  NRT_REQUIRE(!x_quant || x_centers, NRT_E_ARG, "x_quant needs bin centres");
  NRT_REQUIRE(!y_quant || y_centers, NRT_E_ARG, "y_quant needs bin centres");
  NRT_REQUIRE(C == 1 || (x_quant && y_quant), NRT_E_ARG, "channels > 1 only for two quantised operands");
  NRT_REQUIRE(B <= 65535 && C <= 65535, NRT_E_SIZE, "B or C > 65535");
  const int items = B * C;
  NRT_REQUIRE(!dcenters || (workspace && workspace_bytes >= nrt_mi_bwd_workspace_bytes(items)), NRT_E_ARG,
              "centre gradients need the workspace");
  if (nv == 0) return NRT_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BwdArgs a;
  a.x = BwdOperand{x, grad_x, x_batch_stride, x_vox_stride, nbx, x_quant, x_centers};
  a.y = BwdOperand{y, grad_y, y_batch_stride, y_vox_stride, nby, y_quant, y_centers};
  a.nv = nv;
  a.neg_alpha = -alpha;
  a.lo = min_clip;
  a.hi = max_clip;
  a.gstats = gstats;
  a.partial = dcenters ? static_cast<float*>(workspace) : nullptr;
  const int nblk = bwd_blocks(nv, items);
  dim3 grid(nblk, C, B);
  const int want_dcx = dcenters && x_quant && grad_x, want_dcy = dcenters && y_quant && grad_y;
  const int NBP = (nbx <= 16 && nby <= 16) ? 16 : 32;
  if (NBP == 16) launch_bwd<16>(a, grid, want_dcx, want_dcy, st);
  else launch_bwd<32>(a, grid, want_dcx, want_dcy, st);
  int rc = check_launch("mi_bwd_voxel_kernel");
  if (rc != NRT_OK || !dcenters) return rc;
  mi_bwd_combine_kernel<<<1, 128, 0, st>>>(a.partial, (int64_t)nblk * items, NBP, nbx, nby, dcenters, 34);
  return check_launch("mi_bwd_combine_kernel");
}

int nrt_mi_minmax_bwd_f32(const float* x, int64_t n, const float* minmax, const float* dcenters, int nb, float* grad_x,
                          void* stream) {
  NRT_REQUIRE(x && minmax && dcenters && grad_x, NRT_E_ARG, "null pointer");
  NRT_REQUIRE(n >= 0 && nb >= 1 && nb <= 32, NRT_E_ARG, "bad n/nb");
  if (n == 0) return NRT_OK;
  const int grid = (int)imin64((n + 255) / 256, (int64_t)sm_count() * 16);
  mi_minmax_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, minmax, dcenters, nb, grad_x);
  return check_launch("mi_minmax_bwd_kernel");
}

}  // extern "C"
