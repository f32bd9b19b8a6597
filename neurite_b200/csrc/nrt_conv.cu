// nrt_conv.cu -- separable 1-D convolution passes (GaussianBlur) and nearest re-indexing along an
// axis (Subsample): the stages either side of the warp in the synthesis pipeline.
//
// Reference: neurite/tf/utils/utils.py:665-751 separable_conv (tf.nn.convolution = cross-
// correlation of every feature map with a 1-D kernel along one spatial axis, zero 'SAME' padding
// or 'VALID'), utils.py:581-662 gaussian_kernel, layers.py:251-364 GaussianBlur,
// utils.py:754-826 subsample_axis.
//
// A tensor [B, *space, C] seen from axis a is [outer, L, inner] with inner = prod(space[a+1:]) * C
// contiguous.  One pass reads and writes every element once (8 B per element is the roofline
// of a pass); the taps come out of a shared-memory tile:
//   * inner >= 32 ("column" pass): 32 inner elements x 64 outputs per CTA, every thread slides
//     an 8-output register window down the tile (15 shared loads per 64 FMAs);
//   * inner <  32 ("row" pass, e.g. the last axis of a single-channel volume): a contiguous
//     1024-output segment + halo per CTA;
//   * strides / dilations / huge kernels: one thread per output, taps from global memory.
// Accumulation is tap-ascending fp32 FMA; TF's order is unspecified (1e-5 tolerance).
#include "nrt_common.cuh"

namespace nrt {
namespace {

struct ConvArgs {
  const float* x;
  float* out;
  const float* k;      // device [K]
  int64_t outer, L, inner, L_out;
  int K, stride, dil, pad_before;
};

constexpr int kColTL = 64;     // outputs along L per CTA
constexpr int kColRL = 8;      // outputs per thread

// grid.x = outer * l_tiles * i_tiles (i fastest), block (32, 8), smem (kColTL + Kp - 1) * 32 + Kp floats
__global__ void __launch_bounds__(256) sepconv_col_kernel(const ConvArgs a, int Kp, int i_tiles, int l_tiles) {
  extern __shared__ float smem[];
  const int rows = kColTL + Kp - 1;
  float* sm = smem;                 // [rows][32]
  float* ks = smem + rows * 32;     // [Kp], zero padded
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;

  int64_t blk = blockIdx.x;
  const int it = (int)(blk % i_tiles);
  blk /= i_tiles;
  const int lt = (int)(blk % l_tiles);
  const int64_t o = blk / l_tiles;
  const int64_t i = (int64_t)it * 32 + tx;
  const int64_t l0 = (int64_t)lt * kColTL;
  const bool iok = i < a.inner;

  for (int j = tid; j < Kp; j += 256) ks[j] = j < a.K ? a.k[j] : 0.f;
  const float* xo = a.x + o * a.L * a.inner + i;
  for (int r = ty; r < rows; r += 8) {
    const int64_t lg = l0 - a.pad_before + r;
    sm[r * 32 + tx] = (iok && lg >= 0 && lg < a.L) ? ld_stream_f(xo + lg * a.inner) : 0.f;
  }
  __syncthreads();

  float acc[kColRL];
#pragma unroll
  for (int r = 0; r < kColRL; ++r) acc[r] = 0.f;
  const float* col = sm + (ty * kColRL) * 32 + tx;
  float v[kColRL + 7];
#pragma unroll
  for (int u = 0; u < 7; ++u) v[u] = col[u * 32];
  for (int c = 0; c < Kp; c += 8) {
#pragma unroll
    for (int u = 7; u < kColRL + 7; ++u) v[u] = col[(c + u) * 32];
    const float4 k0 = *reinterpret_cast<const float4*>(ks + c);
    const float4 k1 = *reinterpret_cast<const float4*>(ks + c + 4);
    const float kk[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
#pragma unroll
    for (int jj = 0; jj < 8; ++jj)
#pragma unroll
      for (int r = 0; r < kColRL; ++r) acc[r] = fmaf(kk[jj], v[r + jj], acc[r]);
#pragma unroll
    for (int u = 0; u < 7; ++u) v[u] = v[u + 8];
  }
  if (!iok) return;
  float* oo = a.out + o * a.L_out * a.inner + i;
#pragma unroll
  for (int r = 0; r < kColRL; ++r) {
    const int64_t l = l0 + ty * kColRL + r;
    if (l < a.L_out) st_stream_f(oo + l * a.inner, acc[r]);
  }
}

constexpr int kRowTP = 1024;   // output positions per CTA
// grid.x = outer * segs; smem kRowTP + (K-1)*dil*inner + K floats; stride 1 only
__global__ void __launch_bounds__(256) sepconv_row_kernel(const ConvArgs a, int segs) {
  extern __shared__ float smem[];
  const int sp = (int)a.inner * a.dil;                 // tap spacing in elements
  const int halo = (a.K - 1) * sp;
  float* sm = smem;                                    // [kRowTP + halo]
  float* ks = smem + kRowTP + halo;
  const int64_t o = blockIdx.x / segs;
  const int64_t p0 = (int64_t)(blockIdx.x % segs) * kRowTP;
  const int64_t n_in = a.L * a.inner, n_out = a.L_out * a.inner;
  const int64_t src0 = p0 - (int64_t)a.pad_before * a.inner;
  for (int j = threadIdx.x; j < a.K; j += 256) ks[j] = a.k[j];
  const float* xo = a.x + o * n_in;
  for (int e = threadIdx.x; e < kRowTP + halo; e += 256) {
    const int64_t p = src0 + e;
    sm[e] = (p >= 0 && p < n_in) ? ld_stream_f(xo + p) : 0.f;
  }
  __syncthreads();
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const float* base = sm + threadIdx.x;
  for (int j = 0; j < a.K; ++j) {
    const float kj = ks[j];
#pragma unroll
    for (int m = 0; m < 4; ++m) acc[m] = fmaf(kj, base[m * 256 + j * sp], acc[m]);
  }
  float* oo = a.out + o * n_out;
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int64_t p = p0 + threadIdx.x + m * 256;
    if (p < n_out) st_stream_f(oo + p, acc[m]);
  }
}

__global__ void sepconv_generic_kernel(const ConvArgs a) {
  const int64_t total = a.outer * a.L_out * a.inner;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e % a.inner;
    const int64_t t = e / a.inner;
    const int64_t l = t % a.L_out, o = t / a.L_out;
    const float* xo = a.x + o * a.L * a.inner + i;
    float acc = 0.f;
    for (int j = 0; j < a.K; ++j) {
      const int64_t lg = l * a.stride - a.pad_before + (int64_t)j * a.dil;
      if (lg >= 0 && lg < a.L) acc = fmaf(__ldg(a.k + j), __ldg(xo + lg * a.inner), acc);
    }
    a.out[e] = acc;
  }
}

// out[o, l, i] = x[o, idx[l], i]   (tf.gather along an axis, utils.py:818-823)
__global__ void gather_axis_kernel(const float* x, const int32_t* idx, float* out, int64_t outer, int64_t L,
                                   int64_t inner, int64_t L_out) {
  const int64_t total = outer * L_out * inner;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e % inner;
    const int64_t t = e / inner;
    const int64_t l = t % L_out, o = t / L_out;
    const int64_t src = min(max((int64_t)idx[l], (int64_t)0), L - 1);
    out[e] = __ldg(x + (o * L + src) * inner + i);
  }
}

}  // namespace
}  // namespace nrt

using namespace nrt;

extern "C" {

int nrt_sepconv_axis_f32(const float* x, float* out, int64_t outer, int64_t L, int64_t inner, const float* kernel,
                         int K, int stride, int dilation, int pad_before, int64_t L_out, void* stream) {
  NRT_REQUIRE(x && out && kernel, NRT_E_ARG, "null pointer");
  NRT_REQUIRE(x != out, NRT_E_ARG, "in-place convolution is not supported");
  NRT_REQUIRE(outer >= 0 && L >= 1 && inner >= 1 && L_out >= 0, NRT_E_ARG, "bad outer/L/inner/L_out");
  NRT_REQUIRE(K >= 1 && stride >= 1 && dilation >= 1 && pad_before >= 0, NRT_E_ARG, "bad K/stride/dilation/pad");
  const int64_t total = outer * L_out * inner;
  if (total == 0) return NRT_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ConvArgs a{x, out, kernel, outer, L, inner, L_out, K, stride, dilation, pad_before};
  const char* env = getenv("NRT_CONV_GENERIC");
  const bool force_generic = env && atoi(env) != 0;
  const int Kp = (K + 7) & ~7;
  const size_t col_smem = ((size_t)(kColTL + Kp - 1) * 32 + Kp) * sizeof(float);
  const int64_t halo = (int64_t)(K - 1) * dilation * inner;
  const size_t row_smem = (size_t)(kRowTP + halo + K) * sizeof(float);
  if (!force_generic && stride == 1 && dilation == 1 && inner >= 32 && col_smem <= 48 * 1024) {
    const int64_t i_tiles = (inner + 31) / 32, l_tiles = (L_out + kColTL - 1) / kColTL;
    const int64_t nblk = outer * i_tiles * l_tiles;
    NRT_REQUIRE(nblk <= 2147483647LL && i_tiles <= 2147483647LL, NRT_E_SIZE, "tensor too large for one launch");
    sepconv_col_kernel<<<(unsigned)nblk, dim3(32, 8), col_smem, st>>>(a, Kp, (int)i_tiles, (int)l_tiles);
    return check_launch("sepconv_col_kernel");
  }
  if (!force_generic && stride == 1 && inner < 32 && row_smem <= 48 * 1024) {
    const int64_t segs = (L_out * inner + kRowTP - 1) / kRowTP;
    const int64_t nblk = outer * segs;
    NRT_REQUIRE(nblk <= 2147483647LL, NRT_E_SIZE, "tensor too large for one launch");
    sepconv_row_kernel<<<(unsigned)nblk, 256, row_smem, st>>>(a, (int)segs);
    return check_launch("sepconv_row_kernel");
  }
  const int grid = (int)imin64((total + 255) / 256, (int64_t)sm_count() * 32);
  sepconv_generic_kernel<<<grid, 256, 0, st>>>(a);
  return check_launch("sepconv_generic_kernel");
}

int nrt_gather_axis_f32(const float* x, const int32_t* index, float* out, int64_t outer, int64_t L, int64_t inner,
                        int64_t L_out, void* stream) {
  NRT_REQUIRE(x && index && out, NRT_E_ARG, "null pointer");
  NRT_REQUIRE(outer >= 0 && L >= 1 && inner >= 1 && L_out >= 0, NRT_E_ARG, "bad outer/L/inner/L_out");
  const int64_t total = outer * L_out * inner;
  if (total == 0) return NRT_OK;
  const int grid = (int)imin64((total + 255) / 256, (int64_t)sm_count() * 32);
  gather_axis_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, index, out, outer, L, inner, L_out);
  return check_launch("gather_axis_kernel");
}

}  // extern "C"
