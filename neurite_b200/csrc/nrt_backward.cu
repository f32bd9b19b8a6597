// nrt_backward.cu -- gradients of the hot path.  The reference gets
// these from TensorFlow autodiff of the op graph in neurite/tf/utils/utils.py:137-213,
// neurite/tf/metrics.py:471-482 and the Keras cross-entropy; the formulas below are that
// graph differentiated by hand:
//
//   interpn / warp, linear:  out = sum_c w_c * vol[idx_c],  w_c = prod_d w_d(bit_d)
//       d out / d vol[idx_c]  = w_c                                  (scatter-add, tf.gather grad)
//       d out / d loc_d       = [0 <= loc_d <= max_d] * sum_c s_d(c) * prod_{e != d} w_e(c) * vol[idx_c]
//                               with s_d = -1 for corner bit 0 (w = f1 - x) and +1 for bit 1 (w = 1 - (f1 - x));
//                               floor/round have zero gradient, clip_by_value passes it on the closed interval.
//       fill_value: both gradients are multiplied by (1 - oob).
//   nearest: gradient flows to vol only (gather), none to loc.
//   Dice: dice = top/bot (divide_no_nan) or (top+eps)/(bot+eps), top = 2 sum tp, bot = sum t^2 + sum p^2
//       d dice / d p = (2 t (bot+eps) - 2 p (top+eps)) / (bot+eps)^2   (0 where bot == 0 and eps == 0)
//   CCE (from_logits = False): l = -sum_c t_c log clip(p_c / s);  with q = p/s, m_c = [eps <= q_c <= 1-eps]
//       d l / d p_k = -(1/s) * ( t_k m_k / q_k  -  sum_c t_c m_c )      (since q_c / clip(q_c) = 1 where m_c = 1)
#include "nrt_common.cuh"

namespace nrt {

struct BGeo {
  int S[3];
  int C;
  int has_fill;
};

// flat index / weight / sign tables for one point
template <int D>
struct LinSetup {
  int i0[D], i1[D];
  float wlo[D], whi[D];
  bool pass[D];          // clip_by_value passes the gradient (closed interval)
  bool oob;
};

template <int D>
__device__ __forceinline__ void lin_setup(const BGeo& g, const float (&loc)[D], LinSetup<D>& s) {
  s.oob = false;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const float mx = (float)(g.S[d] - 1);
    const Axis a = axis_linear(loc[d], mx, g.S[d] - 1);
    s.i0[d] = a.i0; s.i1[d] = a.i1; s.wlo[d] = a.wlo; s.whi[d] = a.whi;
    s.pass[d] = (loc[d] >= 0.f) && (loc[d] <= mx);
    s.oob = s.oob || (loc[d] < 0.f) || (loc[d] > mx);
  }
}

// one point: scatter grad to vol (atomics), return grad wrt loc
template <int D, int METHOD>
__device__ __forceinline__ void point_backward(const float* __restrict__ vol, float* __restrict__ gvol,
                                               const BGeo& g, const float (&loc)[D],
                                               const float* __restrict__ gout, float (&gloc)[D], bool want_loc) {
#pragma unroll
  for (int d = 0; d < D; ++d) gloc[d] = 0.f;
  if (METHOD == NRT_NEAREST) {
    bool oob = false;
    int idx = 0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      idx = idx * g.S[d] + axis_nearest(loc[d], g.S[d] - 1);
      oob = oob || (loc[d] < 0.f) || (loc[d] > (float)(g.S[d] - 1));
    }
    if (gvol && !(g.has_fill && oob))
      for (int c = 0; c < g.C; ++c) atomicAdd(gvol + (size_t)idx * g.C + c, gout[c]);
    return;
  }
  LinSetup<D> s;
  lin_setup<D>(g, loc, s);
  if (g.has_fill && s.oob) return;                     // out = fill: no dependence on vol or loc
#pragma unroll
  for (int corner = 0; corner < (1 << D); ++corner) {
    int idx = 0;
    float w = 1.f;
    float wex[D];                                       // product of the other axes' weights
#pragma unroll
    for (int d = 0; d < D; ++d) wex[d] = 1.f;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const int bit = (corner >> (D - 1 - d)) & 1;
      idx = idx * g.S[d] + (bit ? s.i1[d] : s.i0[d]);
      const float wd = bit ? s.whi[d] : s.wlo[d];
      w *= wd;
#pragma unroll
      for (int e = 0; e < D; ++e)
        if (e != d) wex[e] *= wd;
    }
    float dot = 0.f;                                    // sum_c gout[c] * vol[idx, c]
    for (int c = 0; c < g.C; ++c) {
      const float go = gout[c];
      if (gvol) atomicAdd(gvol + (size_t)idx * g.C + c, w * go);
      if (want_loc) dot = fmaf(go, __ldg(vol + (size_t)idx * g.C + c), dot);
    }
    if (want_loc) {
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const int bit = (corner >> (D - 1 - d)) & 1;
        gloc[d] += (bit ? dot : -dot) * wex[d];
      }
    }
  }
#pragma unroll
  for (int d = 0; d < D; ++d)
    if (!s.pass[d]) gloc[d] = 0.f;
}

template <int D, int METHOD>
__global__ void __launch_bounds__(256)
warp_bwd_kernel(const float* __restrict__ vol, const float* __restrict__ flow, const float* __restrict__ gout,
                float* __restrict__ gvol, float* __restrict__ gflow, BGeo g, int B, int64_t nvox) {
  const int64_t total = (int64_t)B * nvox;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(t / nvox);
    int rem = (int)(t - (int64_t)b * nvox);
    float loc[D];
    int coord[D];
#pragma unroll
    for (int d = D - 1; d >= 0; --d) { coord[d] = rem % g.S[d]; rem /= g.S[d]; }
#pragma unroll
    for (int d = 0; d < D; ++d) loc[d] = __fadd_rn((float)coord[d], __ldg(flow + t * D + d));
    float gl[D];
    point_backward<D, METHOD>(vol + (size_t)b * nvox * g.C, gvol ? gvol + (size_t)b * nvox * g.C : nullptr, g, loc,
                              gout + t * g.C, gl, gflow != nullptr);
    if (gflow) {
#pragma unroll
      for (int d = 0; d < D; ++d) gflow[t * D + d] = gl[d];
    }
  }
}

template <int D, int METHOD>
__global__ void __launch_bounds__(256)
interpn_bwd_kernel(const float* __restrict__ vol, const float* __restrict__ loc_t, const float* __restrict__ gout,
                   float* __restrict__ gvol, float* __restrict__ gloc_t, BGeo g, int64_t n_out) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_out; t += (int64_t)gridDim.x * blockDim.x) {
    float loc[D], gl[D];
#pragma unroll
    for (int d = 0; d < D; ++d) loc[d] = __ldg(loc_t + t * D + d);
    point_backward<D, METHOD>(vol, gvol, g, loc, gout + t * g.C, gl, gloc_t != nullptr);
    if (gloc_t) {
#pragma unroll
      for (int d = 0; d < D; ++d) gloc_t[t * D + d] = gl[d];
    }
  }
}

template <int D, int METHOD>
__global__ void __launch_bounds__(256)
resize_bwd_kernel(const float* __restrict__ gout, float* __restrict__ gvol, BGeo g, int M0, int M1, int M2,
                  float d0, float d1, float d2, int B, int64_t in_vox, int64_t out_vox) {
  const int M[3] = {M0, M1, M2};
  const float delta[3] = {d0, d1, d2};
  const int64_t total = (int64_t)B * out_vox;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(t / out_vox);
    int rem = (int)(t - (int64_t)b * out_vox);
    float loc[D], gl[D];
#pragma unroll
    for (int d = D - 1; d >= 0; --d) {
      const int i = rem % M[d];
      rem /= M[d];
      loc[d] = (i == M[d] - 1 && M[d] > 1) ? (float)(g.S[d] - 1) : __fmul_rn(delta[d], (float)i);
    }
    point_backward<D, METHOD>(nullptr, gvol + (size_t)b * in_vox * g.C, g, loc, gout + t * g.C, gl, false);
  }
}

// ---------------------------------------------------------------------------------------
// Dice / CCE
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dice_bwd_kernel(const float* __restrict__ t, const float* __restrict__ p, const float* __restrict__ sums,
                const float* __restrict__ gdice, int B, int64_t V, int L, float eps,
                float* __restrict__ gt, float* __restrict__ gp) {
  extern __shared__ float s_coef[];                 // [2][L]: a = 2 G /(bot+eps), c = 2 G (top+eps)/(bot+eps)^2
  const int b = blockIdx.y;
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    const float top = 2.f * sums[((int64_t)b * L + l) * 3 + 0];
    const float bot = sums[((int64_t)b * L + l) * 3 + 1] + sums[((int64_t)b * L + l) * 3 + 2];
    const float G = gdice[(int64_t)b * L + l];
    float a = 0.f, c = 0.f;
    if (eps > 0.f || bot != 0.f) {
      const float den = bot + eps;
      a = 2.f * G / den;
      c = 2.f * G * (top + eps) / (den * den);
    }
    s_coef[l] = a;
    s_coef[L + l] = c;
  }
  __syncthreads();
  const int64_t n = V * L;
  const float* tb = t + (int64_t)b * n;
  const float* pb = p + (int64_t)b * n;
  if ((L & 3) == 0 && (((uintptr_t)t | (uintptr_t)p | (uintptr_t)gt | (uintptr_t)gp) & 15) == 0) {
    // float4 path: a quad never straddles a voxel because L % 4 == 0
    const float4* t4 = reinterpret_cast<const float4*>(tb);
    const float4* p4 = reinterpret_cast<const float4*>(pb);
    const int L4 = L >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n >> 2); i += (int64_t)gridDim.x * blockDim.x) {
      const int l = (int)(i % L4) * 4;
      const float4 tv = ld_stream_f4(t4 + i), pv = ld_stream_f4(p4 + i);
      const float a0 = s_coef[l], a1 = s_coef[l + 1], a2 = s_coef[l + 2], a3 = s_coef[l + 3];
      const float c0 = s_coef[L + l], c1 = s_coef[L + l + 1], c2 = s_coef[L + l + 2], c3 = s_coef[L + l + 3];
      if (gp) st_stream_f4(reinterpret_cast<float4*>(gp + (int64_t)b * n) + i,
                           make_float4(a0 * tv.x - c0 * pv.x, a1 * tv.y - c1 * pv.y, a2 * tv.z - c2 * pv.z, a3 * tv.w - c3 * pv.w));
      if (gt) st_stream_f4(reinterpret_cast<float4*>(gt + (int64_t)b * n) + i,
                           make_float4(a0 * pv.x - c0 * tv.x, a1 * pv.y - c1 * tv.y, a2 * pv.z - c2 * tv.z, a3 * pv.w - c3 * tv.w));
    }
    return;
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int l = (int)(i % L);
    const float tv = ld_stream_f(tb + i), pv = ld_stream_f(pb + i);
    const float a = s_coef[l], c = s_coef[L + l];
    if (gp) gp[(int64_t)b * n + i] = a * tv - c * pv;
    if (gt) gt[(int64_t)b * n + i] = a * pv - c * tv;
  }
}

__global__ void __launch_bounds__(256)
cce_bwd_kernel(const float* __restrict__ t, const float* __restrict__ p, const float* __restrict__ label_w,
               const float* __restrict__ sample_w, int64_t n, int C, int from_logits, float smoothing,
               const float* __restrict__ gscale_ptr, float gscale, const float* __restrict__ gper,
               float* __restrict__ gp) {
  const float eps = 1e-7f, one_m_eps = __fsub_rn(1.0f, 1e-7f);
  const float keep = 1.f - smoothing, add = smoothing / (float)C;
  const float gs = gscale_ptr ? gscale * __ldg(gscale_ptr) : gscale;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    const float* tr = t + r * C;
    const float* pr = p + r * C;
    float up = gs * (sample_w ? __ldg(sample_w + r) : 1.f) * (gper ? __ldg(gper + r) : 1.f);
    if (!from_logits) {
      float s = 0.f;
      for (int c = 0; c < C; ++c) s += pr[c];
      const float rs = 1.f / s;
      float tsum = 0.f;                              // sum_c t_c m_c
      for (int c = 0; c < C; ++c) {
        float tv = tr[c] * (label_w ? __ldg(label_w + c) : 1.f);
        tv = tv * keep + add;
        const float q = pr[c] * rs;
        if (q >= eps && q <= one_m_eps) tsum += tv;
      }
      for (int c = 0; c < C; ++c) {
        float tv = tr[c] * (label_w ? __ldg(label_w + c) : 1.f);
        tv = tv * keep + add;
        const float q = pr[c] * rs;
        const float m = (q >= eps && q <= one_m_eps) ? 1.f : 0.f;
        gp[r * C + c] = -up * rs * ((m != 0.f ? tv / q : 0.f) - tsum);
      }
    } else {
      float mx = pr[0];
      for (int c = 1; c < C; ++c) mx = fmaxf(mx, pr[c]);
      float s = 0.f, tsum = 0.f;
      for (int c = 0; c < C; ++c) s += expf(pr[c] - mx);
      for (int c = 0; c < C; ++c) {
        float tv = tr[c] * (label_w ? __ldg(label_w + c) : 1.f);
        tsum += tv * keep + add;
      }
      for (int c = 0; c < C; ++c) {
        float tv = tr[c] * (label_w ? __ldg(label_w + c) : 1.f);
        tv = tv * keep + add;
        gp[r * C + c] = up * (tsum * expf(pr[c] - mx) / s - tv);     // softmax * sum t - t
      }
    }
  }
}

// vector path (from_logits = False): C/4 lanes per row, coalesced float4 streams
__global__ void __launch_bounds__(256)
cce_bwd_vec4_kernel(const float4* __restrict__ t4, const float4* __restrict__ p4, const float* __restrict__ label_w,
                    const float* __restrict__ sample_w, int64_t n, int C, int q, float smoothing,
                    const float* __restrict__ gscale_ptr, float gscale, const float* __restrict__ gper,
                    float4* __restrict__ gp4) {
  const float eps = 1e-7f, one_m_eps = __fsub_rn(1.0f, 1e-7f);
  const float keep = 1.f - smoothing, add = smoothing / (float)C;
  const float gs = gscale_ptr ? gscale * __ldg(gscale_ptr) : gscale;
  const int tid = threadIdx.x, sub = tid & (q - 1);
  const int rows_per_pass = 256 / q;
  float4 lw = make_float4(1.f, 1.f, 1.f, 1.f);
  if (label_w) lw = __ldg(reinterpret_cast<const float4*>(label_w) + sub);
  for (int64_t r0 = (int64_t)blockIdx.x * rows_per_pass; r0 < n; r0 += (int64_t)gridDim.x * rows_per_pass) {
    const int64_t r = r0 + tid / q;                       // block-uniform trip count (group shuffles below)
    const bool valid = r < n;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f), p = make_float4(1.f, 1.f, 1.f, 1.f);
    if (valid) { t = ld_stream_f4(t4 + r * q + sub); p = ld_stream_f4(p4 + r * q + sub); }
    t.x = t.x * lw.x * keep + add; t.y = t.y * lw.y * keep + add; t.z = t.z * lw.z * keep + add; t.w = t.w * lw.w * keep + add;
    float s = (p.x + p.y) + (p.z + p.w);
    for (int o = q >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float rs = 1.f / s;
    const float q0 = p.x * rs, q1 = p.y * rs, q2 = p.z * rs, q3 = p.w * rs;
    const bool m0 = q0 >= eps && q0 <= one_m_eps, m1 = q1 >= eps && q1 <= one_m_eps;
    const bool m2 = q2 >= eps && q2 <= one_m_eps, m3 = q3 >= eps && q3 <= one_m_eps;
    float ts = ((m0 ? t.x : 0.f) + (m1 ? t.y : 0.f)) + ((m2 ? t.z : 0.f) + (m3 ? t.w : 0.f));
    for (int o = q >> 1; o > 0; o >>= 1) ts += __shfl_xor_sync(0xffffffffu, ts, o);
    if (valid) {
      const float up = -gs * (sample_w ? __ldg(sample_w + r) : 1.f) * (gper ? __ldg(gper + r) : 1.f) * rs;
      float4 g;
      g.x = up * ((m0 ? t.x / q0 : 0.f) - ts);
      g.y = up * ((m1 ? t.y / q1 : 0.f) - ts);
      g.z = up * ((m2 ? t.z / q2 : 0.f) - ts);
      g.w = up * ((m3 ? t.w / q3 : 0.f) - ts);
      gp4[r * q + sub] = g;
    }
  }
}

static int grid_for(int64_t n) { return (int)imin64((n + 255) / 256, (int64_t)sm_count() * 32); }

#define NRT_BWD_DISPATCH(D, method, CALL)                                       \
  do {                                                                          \
    if ((method) == NRT_LINEAR) {                                               \
      if ((D) == 1) { CALL(1, NRT_LINEAR); } else if ((D) == 2) { CALL(2, NRT_LINEAR); } else { CALL(3, NRT_LINEAR); } \
    } else {                                                                    \
      if ((D) == 1) { CALL(1, NRT_NEAREST); } else if ((D) == 2) { CALL(2, NRT_NEAREST); } else { CALL(3, NRT_NEAREST); } \
    }                                                                           \
  } while (0)

static int fill_geo(BGeo& g, const int32_t* shape, int D, int C, int has_fill, int64_t* nvox) {
  NRT_REQUIRE(D >= 1 && D <= 3 && C >= 1 && shape, NRT_E_ARG, "bad D/C/shape");
  *nvox = 1;
  for (int d = 0; d < 3; ++d) {
    g.S[d] = d < D ? shape[d] : 1;
    NRT_REQUIRE(g.S[d] >= 1, NRT_E_ARG, "shape[%d] = %d", d, g.S[d]);
    *nvox *= g.S[d];
  }
  NRT_REQUIRE(*nvox <= 0x7fffffffLL, NRT_E_SIZE, "volume too large for int32 indexing");
  g.C = C; g.has_fill = has_fill;
  return NRT_OK;
}

}  // namespace nrt

using namespace nrt;

extern "C" {

int nrt_warp_bwd_f32(const float* vol, const float* flow, const float* grad_out, float* grad_vol, float* grad_flow,
                     int B, const int32_t* shape, int D, int C, int method, int has_fill, void* stream) {
  NRT_REQUIRE(vol && flow && grad_out && (grad_vol || grad_flow), NRT_E_ARG, "null pointer");
  NRT_REQUIRE(method == NRT_LINEAR || method == NRT_NEAREST, NRT_E_ARG, "method should be linear or nearest, got: %d", method);
  BGeo g;
  int64_t nvox;
  int rc = fill_geo(g, shape, D, C, has_fill, &nvox);
  if (rc != NRT_OK) return rc;
  if (B <= 0 || nvox == 0) return NRT_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (method == NRT_NEAREST && grad_flow) cudaMemsetAsync(grad_flow, 0, (size_t)B * nvox * D * sizeof(float), st);
  float* gf = method == NRT_NEAREST ? nullptr : grad_flow;
  if (D == 3 && C == 1 && (grad_vol || gf)) {
    bool used = false;
    rc = warp3d_bwd_tile(vol, flow, grad_out, grad_vol, gf, B, shape, method, has_fill, st, &used);
    if (rc != NRT_OK || used) return rc;
  }
  if (!grad_vol && !gf) return check_launch("warp_bwd memset");
#define CALL(DD, MM) warp_bwd_kernel<DD, MM><<<grid_for((int64_t)B * nvox), 256, 0, st>>>(vol, flow, grad_out, grad_vol, gf, g, B, nvox)
  NRT_BWD_DISPATCH(D, method, CALL);
#undef CALL
  return check_launch("warp_bwd_kernel");
}

int nrt_interpn_bwd_f32(const float* vol, const int32_t* vol_shape, int D, int C, const float* loc, int64_t n_out,
                        int method, int has_fill, const float* grad_out, float* grad_vol, float* grad_loc, void* stream) {
  NRT_REQUIRE(vol && loc && grad_out && (grad_vol || grad_loc), NRT_E_ARG, "null pointer");
  NRT_REQUIRE(method == NRT_LINEAR || method == NRT_NEAREST, NRT_E_ARG, "method should be linear or nearest, got: %d", method);
  BGeo g;
  int64_t nvox;
  int rc = fill_geo(g, vol_shape, D, C, has_fill, &nvox);
  if (rc != NRT_OK) return rc;
  if (n_out <= 0) return NRT_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (method == NRT_NEAREST && grad_loc) cudaMemsetAsync(grad_loc, 0, (size_t)n_out * D * sizeof(float), st);
  float* gl = method == NRT_NEAREST ? nullptr : grad_loc;
  if (!grad_vol && !gl) return check_launch("interpn_bwd memset");
#define CALL(DD, MM) interpn_bwd_kernel<DD, MM><<<grid_for(n_out), 256, 0, st>>>(vol, loc, grad_out, grad_vol, gl, g, n_out)
  NRT_BWD_DISPATCH(D, method, CALL);
#undef CALL
  return check_launch("interpn_bwd_kernel");
}

int nrt_resize_bwd_f32(const float* grad_out, float* grad_vol, int B, const int32_t* in_shape, const int32_t* out_shape,
                       int D, int C, int method, void* stream) {
  NRT_REQUIRE(grad_out && grad_vol && out_shape, NRT_E_ARG, "null pointer");
  NRT_REQUIRE(method == NRT_LINEAR || method == NRT_NEAREST, NRT_E_ARG, "method should be linear or nearest, got: %d", method);
  BGeo g;
  int64_t in_vox;
  int rc = fill_geo(g, in_shape, D, C, 0, &in_vox);
  if (rc != NRT_OK) return rc;
  int M[3] = {1, 1, 1};
  float delta[3] = {0.f, 0.f, 0.f};
  int64_t out_vox = 1;
  for (int d = 0; d < D; ++d) {
    M[d] = out_shape[d];
    NRT_REQUIRE(M[d] >= 0, NRT_E_ARG, "out_shape[%d] = %d", d, M[d]);
    delta[d] = M[d] > 1 ? (float)(g.S[d] - 1) / (float)(M[d] - 1) : 0.0f;
    out_vox *= M[d];
  }
  NRT_REQUIRE(out_vox <= 0x7fffffffLL, NRT_E_SIZE, "output too large for int32 indexing");
  if (B <= 0 || out_vox == 0) return NRT_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define CALL(DD, MM) resize_bwd_kernel<DD, MM><<<grid_for((int64_t)B * out_vox), 256, 0, st>>>(grad_out, grad_vol, g, M[0], M[1], M[2], delta[0], delta[1], delta[2], B, in_vox, out_vox)
  NRT_BWD_DISPATCH(D, method, CALL);
#undef CALL
  return check_launch("resize_bwd_kernel");
}

int nrt_dice_bwd_f32(const float* y_true, const float* y_pred, const float* sums, const float* grad_dice, int B,
                     int64_t V, int L, float laplace, float* grad_true, float* grad_pred, void* stream) {
  NRT_REQUIRE(y_true && y_pred && sums && grad_dice && (grad_true || grad_pred), NRT_E_ARG, "null pointer");
  NRT_REQUIRE(B >= 1 && B <= 65535 && L >= 1 && V >= 0, NRT_E_ARG, "bad B/L/V");
  NRT_REQUIRE((size_t)2 * L * sizeof(float) <= 48 * 1024, NRT_E_SIZE, "L = %d too large", L);
  if (V == 0) return NRT_OK;
  int gx = (int)imin64((V * L + 1023) / 1024, (int64_t)(sm_count() * 16 + B - 1) / B);
  if (gx < 1) gx = 1;
  dim3 grid(gx, B);
  dice_bwd_kernel<<<grid, 256, 2 * L * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      y_true, y_pred, sums, grad_dice, B, V, L, laplace, grad_true, grad_pred);
  return check_launch("dice_bwd_kernel");
}

int nrt_cce_bwd_f32(const float* y_true, const float* y_pred, const float* label_w, const float* sample_w, int64_t n,
                    int C, int from_logits, float label_smoothing, const float* grad_scalar, float scale,
                    const float* grad_per_elem, float* grad_pred, void* stream) {
  NRT_REQUIRE(y_true && y_pred && grad_pred, NRT_E_ARG, "null pointer");
  NRT_REQUIRE(n >= 0 && C >= 1, NRT_E_ARG, "bad n/C");
  if (n == 0) return NRT_OK;
  const int q = C / 4;
  if (!from_logits && C % 4 == 0 && q <= 32 && (q & (q - 1)) == 0 && aligned16(y_true) && aligned16(y_pred) &&
      aligned16(grad_pred) && (!label_w || aligned16(label_w))) {
    const int rows_per_pass = 256 / q;
    const int grid = (int)imin64((n + rows_per_pass - 1) / rows_per_pass, (int64_t)sm_count() * 16);
    cce_bwd_vec4_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(y_true), reinterpret_cast<const float4*>(y_pred), label_w, sample_w, n, C, q,
        label_smoothing, grad_scalar, scale, grad_per_elem, reinterpret_cast<float4*>(grad_pred));
    return check_launch("cce_bwd_vec4_kernel");
  }
  cce_bwd_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      y_true, y_pred, label_w, sample_w, n, C, from_logits, label_smoothing, grad_scalar, scale, grad_per_elem, grad_pred);
  return check_launch("cce_bwd_kernel");
}

}  // extern "C"
