// nrt_interp.cu -- N-D gridded interpolation for sm_100a:
//   nrt_interpn_f32  (explicit loc tensor)      reference utils.py:73-220
//   nrt_warp_f32     (identity grid + flow)     voxelmorph SpatialTransformer contract
//   nrt_resize_f32   (in-kernel linspace grid)  reference utils.py:223-265, layers.py:154-181
//
// Two kernel families:
//   * generic gather kernels (any D in 1..3, any C, linear/nearest): one thread per
//     (output point, 4-channel chunk); corners read through L1/L2 with __ldg.
//   * warp3d_tile_kernel (D=3, C=1): the hot kernel of BASELINE.json.  One CTA per
//     TZxTYx32 output tile; the flow tile and the bounding source box (tile + halo) are
//     staged into shared memory by two TMA tensor loads completing on one mbarrier; the
//     8-corner gather runs from shared memory (bank-conflict bound instead of L1-sector
//     bound for incoherent flows).  Any corner outside the staged box falls back, per
//     voxel, to the global gather, so results never depend on the halo.
//
// Arithmetic is bit-faithful to the reference's unfused TF ops: every multiply and add is
// a separate fp32 rounding (__fmul_rn/__fadd_rn), corner order is itertools.product order.
#include <cuda.h>   // CUtensorMap (types only; the encode entry point is fetched at run time)

#include "nrt_common.cuh"

namespace nrt {

struct Geo {
  int S[3];        // full spatial extent of the source volume per axis (clip bounds)
  int src_z0;      // global index of the first resident source plane (axis 0)
  int src_n0;      // resident source planes
  int C;
  int has_fill;
  float fill;
  int32_t* err;    // device flag: corner outside the resident planes
};

// flat row-major index over the RESIDENT source (src_n0, S1, S2), reference sub2ind2d
template <int D>
__device__ __forceinline__ int flat_index(const Geo& g, const int (&sub)[D]) {
  int ndx = sub[0];
#pragma unroll
  for (int d = 1; d < D; ++d) ndx = ndx * g.S[d] + sub[d];
  return ndx;
}

__device__ __forceinline__ int to_resident(const Geo& g, int i) {
  int l = i - g.src_z0;
  if (l < 0 || l >= g.src_n0) {
    if (g.err) atomicOr(g.err, 1);
    l = min(max(l, 0), g.src_n0 - 1);
  }
  return l;
}

template <int D>
__device__ __forceinline__ bool out_of_bounds(const Geo& g, const float (&loc)[D]) {
  bool oob = false;
#pragma unroll
  for (int d = 0; d < D; ++d) oob = oob || (loc[d] < 0.0f) || (loc[d] > (float)(g.S[d] - 1));
  return oob;
}

// One output point, channels [c0, c0+VEC) -- generic global-memory gather.
template <int D, int VEC, int METHOD>
__device__ __forceinline__ void sample_point(const float* __restrict__ vol, const Geo& g,
                                             const float (&loc)[D], int c0, float (&res)[VEC]) {
  constexpr int NC = 1 << D;
  if (METHOD == NRT_LINEAR) {
    Axis a[D];
#pragma unroll
    for (int d = 0; d < D; ++d) a[d] = axis_linear(loc[d], (float)(g.S[d] - 1), g.S[d] - 1);
    a[0].i0 = to_resident(g, a[0].i0);
    a[0].i1 = to_resident(g, a[0].i1);
#pragma unroll
    for (int v = 0; v < VEC; ++v) res[v] = 0.0f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      int sub[D];
      float w = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const int bit = (c >> (D - 1 - d)) & 1;          // first axis = most significant
        sub[d] = bit ? a[d].i1 : a[d].i0;
        const float wd = bit ? a[d].whi : a[d].wlo;
        w = (d == 0) ? wd : __fmul_rn(w, wd);             // prod_n: ((w0*w1)*w2)
      }
      const size_t off = (size_t)flat_index<D>(g, sub) * g.C + c0;
      if (VEC == 4) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(vol + off));
        res[0] = __fadd_rn(res[0], __fmul_rn(w, q.x));
        res[1 % VEC] = __fadd_rn(res[1 % VEC], __fmul_rn(w, q.y));
        res[2 % VEC] = __fadd_rn(res[2 % VEC], __fmul_rn(w, q.z));
        res[3 % VEC] = __fadd_rn(res[3 % VEC], __fmul_rn(w, q.w));
      } else {
        res[0] = __fadd_rn(res[0], __fmul_rn(w, __ldg(vol + off)));
      }
    }
  } else {
    int sub[D];
#pragma unroll
    for (int d = 0; d < D; ++d) sub[d] = axis_nearest(loc[d], g.S[d] - 1);
    sub[0] = to_resident(g, sub[0]);
    const size_t off = (size_t)flat_index<D>(g, sub) * g.C + c0;
    if (VEC == 4) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(vol + off));
      res[0] = q.x; res[1 % VEC] = q.y; res[2 % VEC] = q.z; res[3 % VEC] = q.w;
    } else {
      res[0] = __ldg(vol + off);
    }
  }
  if (g.has_fill) {
    const bool oob = out_of_bounds<D>(g, loc);
#pragma unroll
    for (int v = 0; v < VEC; ++v) res[v] = apply_fill(res[v], oob, g.fill);
  }
}

template <int VEC>
__device__ __forceinline__ void store_vec(float* __restrict__ p, const float (&r)[VEC]) {
  if (VEC == 4) *reinterpret_cast<float4*>(p) = make_float4(r[0], r[1 % VEC], r[2 % VEC], r[3 % VEC]);
  else p[0] = r[0];
}

// ---------------------------------------------------------------------------------------
// generic kernels
// ---------------------------------------------------------------------------------------
template <int D, int VEC, int METHOD>
__global__ void __launch_bounds__(256)
interpn_kernel(const float* __restrict__ vol, Geo g, const float* __restrict__ loc_t,
               int64_t n_out, float* __restrict__ out) {
  const int cv_n = g.C / VEC;
  const int64_t total = n_out * cv_n;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pt = t / cv_n;
    const int c0 = (int)(t - pt * cv_n) * VEC;
    float loc[D];
#pragma unroll
    for (int d = 0; d < D; ++d) loc[d] = __ldg(loc_t + pt * D + d);
    float r[VEC];
    sample_point<D, VEC, METHOD>(vol, g, loc, c0, r);
    store_vec<VEC>(out + pt * g.C + c0, r);
  }
}

struct WarpGeo {
  Geo g;
  int out_z0, out_n0;   // produced planes of axis 0
  int64_t src_batch_stride, out_vox;   // elements per batch item of vol; voxels per batch item of out
  int B;
};

template <int D, int VEC, int METHOD>
__global__ void __launch_bounds__(256)
warp_generic_kernel(const float* __restrict__ vol, const float* __restrict__ flow,
                    float* __restrict__ out, WarpGeo w) {
  const Geo& g = w.g;
  const int cv_n = g.C / VEC;
  const int64_t total = (int64_t)w.B * w.out_vox * cv_n;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pv = t / cv_n;                       // (b, voxel)
    const int c0 = (int)(t - pv * cv_n) * VEC;
    const int b = (int)(pv / w.out_vox);
    int rem = (int)(pv - (int64_t)b * w.out_vox);
    int coord[D];
#pragma unroll
    for (int d = D - 1; d >= 1; --d) { coord[d] = rem % g.S[d]; rem /= g.S[d]; }
    coord[0] = rem + w.out_z0;
    float loc[D];
#pragma unroll
    for (int d = 0; d < D; ++d) loc[d] = __fadd_rn((float)coord[d], __ldg(flow + pv * D + d));
    float r[VEC];
    sample_point<D, VEC, METHOD>(vol + (size_t)b * w.src_batch_stride, g, loc, c0, r);
    store_vec<VEC>(out + pv * g.C + c0, r);
  }
}

struct ResizeGeo {
  Geo g;
  int M[3];          // full output extent
  float delta[3];    // fp32 (S-1)/(M-1)
  int out_z0, out_n0;
  int64_t src_batch_stride, out_vox;
  int B;
};

template <int D, int VEC, int METHOD>
__global__ void __launch_bounds__(256)
resize_kernel(const float* __restrict__ vol, float* __restrict__ out, ResizeGeo w) {
  const Geo& g = w.g;
  const int cv_n = g.C / VEC;
  const int64_t total = (int64_t)w.B * w.out_vox * cv_n;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pv = t / cv_n;
    const int c0 = (int)(t - pv * cv_n) * VEC;
    const int b = (int)(pv / w.out_vox);
    int rem = (int)(pv - (int64_t)b * w.out_vox);
    int coord[D];
#pragma unroll
    for (int d = D - 1; d >= 1; --d) { coord[d] = rem % w.M[d]; rem /= w.M[d]; }
    coord[0] = rem + w.out_z0;
    float loc[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      // tf.linspace(0, S-1, M): endpoints exact, interior 0 + delta*i  (utils.py:259)
      const int i = coord[d];
      loc[d] = (i == w.M[d] - 1 && w.M[d] > 1) ? (float)(g.S[d] - 1) : __fmul_rn(w.delta[d], (float)i);
    }
    float r[VEC];
    sample_point<D, VEC, METHOD>(vol + (size_t)b * w.src_batch_stride, g, loc, c0, r);
    store_vec<VEC>(out + pv * g.C + c0, r);
  }
}

// ---------------------------------------------------------------------------------------
// warp3d_tile_kernel: D=3, C=1, TMA-staged flow tile + source box in shared memory
// ---------------------------------------------------------------------------------------
struct TileGeo {
  Geo g;                 // S = {full_s0, H, W}
  int out_z0, out_n0;
  int B;
  int BZ, BY, BX;        // source box extent (elements)
  int hz, hy, hx;        // box origin = tile origin - h
  int ntz, nty, ntx;     // tiles per axis
  int64_t src_batch_stride, out_vox;
};

template <int TZ, int TY, int METHOD>
__global__ void __launch_bounds__(256)
warp3d_tile_kernel(const __grid_constant__ CUtensorMap tm_vol,
                   const __grid_constant__ CUtensorMap tm_flow,
                   const float* __restrict__ vol, float* __restrict__ out, TileGeo w) {
  constexpr int TX = 32;
  constexpr int NW = 8;                       // warps per CTA; warp = one x-row of 32 voxels
  static_assert(TY % NW == 0, "TY must be a multiple of the warp count");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* s_flow = reinterpret_cast<float*>(smem_raw);                       // [TZ][TY][TX][3]
  float* s_box = s_flow + TZ * TY * TX * 3;                                 // [BZ][BY][BX]
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_box + w.BZ * w.BY * w.BX);

  const Geo& g = w.g;
  int tile = blockIdx.x;
  const int tx = tile % w.ntx; tile /= w.ntx;
  const int ty = tile % w.nty; tile /= w.nty;
  const int tz = tile % w.ntz;
  const int b = tile / w.ntz;
  const int x0 = tx * TX, y0 = ty * TY, z0l = tz * TZ;      // z0l: plane within the output slab
  const int gz0 = w.out_z0 + z0l;                           // global z of the tile's first plane
  const int ox = x0 - w.hx, oy = y0 - w.hy, oz = gz0 - w.hz;   // global coords of box origin

  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
    const uint32_t bytes = (uint32_t)((TZ * TY * TX * 3 + w.BZ * w.BY * w.BX) * sizeof(float));
    mbar_expect_tx(bar, bytes);
    tma_load_4d(s_flow, &tm_flow, bar, x0 * 3, y0, z0l, b);
    tma_load_4d(s_box, &tm_vol, bar, ox, oy, oz - g.src_z0, b);
  }
  __syncthreads();

  // region of the box that holds real (resident, in-volume) voxels, in global coordinates
  const int lo_z = max(oz, g.src_z0), hi_z = min(oz + w.BZ - 1, g.src_z0 + g.src_n0 - 1);
  const int lo_y = max(oy, 0), hi_y = min(oy + w.BY - 1, g.S[1] - 1);
  const int lo_x = max(ox, 0), hi_x = min(ox + w.BX - 1, g.S[2] - 1);
  const float mz = (float)(g.S[0] - 1), my = (float)(g.S[1] - 1), mx = (float)(g.S[2] - 1);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int gx = x0 + lane;
  const float fx = (float)gx;
  const bool x_ok = gx < g.S[2];
  const float* volb = vol + (size_t)b * w.src_batch_stride;
  float* outb = out + (size_t)b * w.out_vox;
  const int sBY = w.BY, sBX = w.BX;
  const int box_base = -((oz * sBY + oy) * sBX + ox);

  mbar_wait(bar, 0);

#pragma unroll 1
  for (int z = 0; z < TZ; ++z) {
    const int gz = gz0 + z;
    const float fz = (float)gz;
#pragma unroll 2
    for (int yy = wid; yy < TY; yy += NW) {
      const int gy = y0 + yy;
      const float* fl = s_flow + ((z * TY + yy) * TX + lane) * 3;
      const float lz = __fadd_rn(fz, fl[0]);
      const float ly = __fadd_rn((float)gy, fl[1]);
      const float lx = __fadd_rn(fx, fl[2]);
      float res;
      if (METHOD == NRT_LINEAR) {
        const float cz = fminf(fmaxf(lz, 0.f), mz), cy = fminf(fmaxf(ly, 0.f), my), cx = fminf(fmaxf(lx, 0.f), mx);
        const float f0z = floorf(cz), f0y = floorf(cy), f0x = floorf(cx);
        const int iz = __float2int_rz(f0z), iy = __float2int_rz(f0y), ix = __float2int_rz(f0x);
        float v[8];
        float wz0, wz1, wy0, wy1, wx0, wx1;
        const bool fast = (iz >= lo_z) & (iz < hi_z) & (iy >= lo_y) & (iy < hi_y) & (ix >= lo_x) & (ix < hi_x);
        if (fast) {
          // i1 = i0 + 1 needs no clip here; weights exactly as axis_linear computes them
          wz0 = __fsub_rn(__fadd_rn(f0z, 1.f), cz); wz1 = __fsub_rn(1.f, wz0);
          wy0 = __fsub_rn(__fadd_rn(f0y, 1.f), cy); wy1 = __fsub_rn(1.f, wy0);
          wx0 = __fsub_rn(__fadd_rn(f0x, 1.f), cx); wx1 = __fsub_rn(1.f, wx0);
          const float* p = s_box + (box_base + (iz * sBY + iy) * sBX + ix);
          const int dy = sBX, dz = sBY * sBX;
          v[0] = p[0];       v[1] = p[1];
          v[2] = p[dy];      v[3] = p[dy + 1];
          v[4] = p[dz];      v[5] = p[dz + 1];
          v[6] = p[dz + dy]; v[7] = p[dz + dy + 1];
        } else {
          Axis az = axis_linear(lz, mz, g.S[0] - 1);
          const Axis ay = axis_linear(ly, my, g.S[1] - 1);
          const Axis ax = axis_linear(lx, mx, g.S[2] - 1);
          az.i0 = to_resident(g, az.i0);
          az.i1 = to_resident(g, az.i1);
          wz0 = az.wlo; wz1 = az.whi; wy0 = ay.wlo; wy1 = ay.whi; wx0 = ax.wlo; wx1 = ax.whi;
          const int r00 = (az.i0 * g.S[1] + ay.i0) * g.S[2], r01 = (az.i0 * g.S[1] + ay.i1) * g.S[2];
          const int r10 = (az.i1 * g.S[1] + ay.i0) * g.S[2], r11 = (az.i1 * g.S[1] + ay.i1) * g.S[2];
          v[0] = __ldg(volb + r00 + ax.i0); v[1] = __ldg(volb + r00 + ax.i1);
          v[2] = __ldg(volb + r01 + ax.i0); v[3] = __ldg(volb + r01 + ax.i1);
          v[4] = __ldg(volb + r10 + ax.i0); v[5] = __ldg(volb + r10 + ax.i1);
          v[6] = __ldg(volb + r11 + ax.i0); v[7] = __ldg(volb + r11 + ax.i1);
        }
        const float w00 = __fmul_rn(wz0, wy0), w01 = __fmul_rn(wz0, wy1);
        const float w10 = __fmul_rn(wz1, wy0), w11 = __fmul_rn(wz1, wy1);
        res = __fadd_rn(0.f, __fmul_rn(__fmul_rn(w00, wx0), v[0]));
        res = __fadd_rn(res, __fmul_rn(__fmul_rn(w00, wx1), v[1]));
        res = __fadd_rn(res, __fmul_rn(__fmul_rn(w01, wx0), v[2]));
        res = __fadd_rn(res, __fmul_rn(__fmul_rn(w01, wx1), v[3]));
        res = __fadd_rn(res, __fmul_rn(__fmul_rn(w10, wx0), v[4]));
        res = __fadd_rn(res, __fmul_rn(__fmul_rn(w10, wx1), v[5]));
        res = __fadd_rn(res, __fmul_rn(__fmul_rn(w11, wx0), v[6]));
        res = __fadd_rn(res, __fmul_rn(__fmul_rn(w11, wx1), v[7]));
      } else {
        const int iz = axis_nearest(lz, g.S[0] - 1), iy = axis_nearest(ly, g.S[1] - 1), ix = axis_nearest(lx, g.S[2] - 1);
        const bool fast = (iz >= lo_z) & (iz <= hi_z) & (iy >= lo_y) & (iy <= hi_y) & (ix >= lo_x) & (ix <= hi_x);
        if (fast) res = s_box[box_base + (iz * sBY + iy) * sBX + ix];
        else res = __ldg(volb + (to_resident(g, iz) * g.S[1] + iy) * g.S[2] + ix);
      }
      if (g.has_fill) {
        const bool oob = (lz < 0.f) | (lz > mz) | (ly < 0.f) | (ly > my) | (lx < 0.f) | (lx > mx);
        res = apply_fill(res, oob, g.fill);
      }
      if (x_ok && gy < g.S[1] && (z0l + z) < w.out_n0)
        outb[((size_t)(z0l + z) * g.S[1] + gy) * g.S[2] + gx] = res;
    }
  }
}

// ---------------------------------------------------------------------------------------
// host: tensor-map encoding through the runtime's driver entry point (no -lcuda needed)
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    (void)cudaGetLastError();
  }
  return fn;
}

static int encode_f32_4d(CUtensorMap* tm, const void* base, const uint64_t dims[4], const uint32_t box[4]) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_error(NRT_E_NODEV, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[4], gstr[3];
  cuuint32_t bx[4], es[4] = {1, 1, 1, 1};
  for (int i = 0; i < 4; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; }
  gstr[0] = dims[0] * sizeof(float);
  gstr[1] = gstr[0] * dims[1];
  gstr[2] = gstr[1] * dims[2];
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(NRT_E_LAUNCH, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
  return NRT_OK;
}

static int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return (s && *s) ? atoi(s) : dflt;
}

template <int TZ, int TY, int METHOD>
static int launch_tile(const CUtensorMap& tmv, const CUtensorMap& tmf, const float* vol, float* out,
                       const TileGeo& tg, size_t smem, cudaStream_t st) {
  auto kern = warp3d_tile_kernel<TZ, TY, METHOD>;
  static size_t configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return check_launch("cudaFuncSetAttribute(warp3d_tile)");
    configured = smem;
  }
  const int grid = tg.B * tg.ntz * tg.nty * tg.ntx;
  kern<<<grid, 256, smem, st>>>(tmv, tmf, vol, out, tg);
  return check_launch("warp3d_tile_kernel");
}

template <int D, int METHOD>
static int launch_warp_generic(const float* vol, const float* flow, float* out, const WarpGeo& wg, cudaStream_t st) {
  const bool vec = (wg.g.C % 4 == 0) && aligned16(vol) && aligned16(out);
  const int64_t total = (int64_t)wg.B * wg.out_vox * (vec ? wg.g.C / 4 : wg.g.C);
  if (total == 0) return NRT_OK;
  const int grid = (int)imin64((total + 255) / 256, (int64_t)sm_count() * 32);
  if (vec) warp_generic_kernel<D, 4, METHOD><<<grid, 256, 0, st>>>(vol, flow, out, wg);
  else warp_generic_kernel<D, 1, METHOD><<<grid, 256, 0, st>>>(vol, flow, out, wg);
  return check_launch("warp_generic_kernel");
}

template <int D, int METHOD>
static int launch_interpn(const float* vol, const Geo& g, const float* loc, int64_t n_out, float* out, cudaStream_t st) {
  const bool vec = (g.C % 4 == 0) && aligned16(vol) && aligned16(out);
  const int64_t total = n_out * (vec ? g.C / 4 : g.C);
  if (total == 0) return NRT_OK;
  const int grid = (int)imin64((total + 255) / 256, (int64_t)sm_count() * 32);
  if (vec) interpn_kernel<D, 4, METHOD><<<grid, 256, 0, st>>>(vol, g, loc, n_out, out);
  else interpn_kernel<D, 1, METHOD><<<grid, 256, 0, st>>>(vol, g, loc, n_out, out);
  return check_launch("interpn_kernel");
}

template <int D, int METHOD>
static int launch_resize(const float* vol, float* out, const ResizeGeo& rg, cudaStream_t st) {
  const bool vec = (rg.g.C % 4 == 0) && aligned16(vol) && aligned16(out);
  const int64_t total = (int64_t)rg.B * rg.out_vox * (vec ? rg.g.C / 4 : rg.g.C);
  if (total == 0) return NRT_OK;
  const int grid = (int)imin64((total + 255) / 256, (int64_t)sm_count() * 32);
  if (vec) resize_kernel<D, 4, METHOD><<<grid, 256, 0, st>>>(vol, out, rg);
  else resize_kernel<D, 1, METHOD><<<grid, 256, 0, st>>>(vol, out, rg);
  return check_launch("resize_kernel");
}

#define NRT_DISPATCH_D_METHOD(D, method, CALL)                                  \
  do {                                                                          \
    if ((method) == NRT_LINEAR) {                                               \
      if ((D) == 1) return CALL(1, NRT_LINEAR);                                 \
      if ((D) == 2) return CALL(2, NRT_LINEAR);                                 \
      return CALL(3, NRT_LINEAR);                                               \
    } else {                                                                    \
      if ((D) == 1) return CALL(1, NRT_NEAREST);                                \
      if ((D) == 2) return CALL(2, NRT_NEAREST);                                \
      return CALL(3, NRT_NEAREST);                                              \
    }                                                                           \
  } while (0)

static int check_common(int D, int C, int method) {
  NRT_REQUIRE(D >= 1 && D <= 3, NRT_E_ARG, "D must be 1, 2 or 3 (got %d)", D);
  NRT_REQUIRE(C >= 1, NRT_E_ARG, "C must be >= 1 (got %d)", C);
  NRT_REQUIRE(method == NRT_LINEAR || method == NRT_NEAREST, NRT_E_ARG,
              "method should be linear or nearest, got: %d", method);
  return NRT_OK;
}

static int try_tile_path(const float* vol, const float* flow, float* out, int B, const int32_t* shape,
                         int method, int has_fill, float fill, int src_z0, int src_n0, int out_z0,
                         int out_n0, int halo, int32_t* err_flag, cudaStream_t st, bool* used) {
  *used = false;
  const int H = shape[1], W = shape[2];
  if (env_int("NRT_WARP_TILE", 1) == 0) return NRT_OK;
  if (W % 4 != 0 || !aligned16(vol) || !aligned16(flow) || W < 32) return NRT_OK;
  int cfg = env_int("NRT_WARP_TILE_CFG", 0);         // 0: 8x16x32, 1: 16x16x32, 2: 8x8x32, 3: 4x8x32
  const int TZs[4] = {8, 16, 8, 4}, TYs[4] = {16, 16, 8, 8};
  if (cfg < 0 || cfg > 3) cfg = 0;
  if (halo <= 0) halo = 3;
  const int TX = 32;
  TileGeo tg;
  int TZ, TY;
  size_t smem;
  for (;;) {        // shrink the halo until tile + box fit in shared memory
    TZ = TZs[cfg]; TY = TYs[cfg];
    tg.hz = halo; tg.hy = halo; tg.hx = halo;
    tg.BZ = TZ + 2 * halo; tg.BY = TY + 2 * halo;
    tg.BX = (TX + 2 * halo + 3) & ~3;
    smem = (size_t)(TZ * TY * TX * 3 + tg.BZ * tg.BY * tg.BX) * sizeof(float) + 16;
    if ((smem <= 227 * 1024 && tg.BX <= 256 && tg.BY <= 256 && tg.BZ <= 256) || halo == 1) break;
    --halo;
  }
  if (smem > 227 * 1024) return NRT_OK;
  tg.g.S[0] = shape[0]; tg.g.S[1] = H; tg.g.S[2] = W;
  tg.g.src_z0 = src_z0; tg.g.src_n0 = src_n0; tg.g.C = 1;
  tg.g.has_fill = has_fill; tg.g.fill = fill; tg.g.err = err_flag;
  tg.out_z0 = out_z0; tg.out_n0 = out_n0; tg.B = B;
  tg.ntz = (out_n0 + TZ - 1) / TZ; tg.nty = (H + TY - 1) / TY; tg.ntx = (W + TX - 1) / TX;
  tg.src_batch_stride = (int64_t)src_n0 * H * W;
  tg.out_vox = (int64_t)out_n0 * H * W;
  if ((int64_t)tg.B * tg.ntz * tg.nty * tg.ntx > 0x7fffffffLL) return NRT_OK;

  CUtensorMap tmv, tmf;
  const uint64_t vd[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)src_n0, (uint64_t)B};
  const uint32_t vb[4] = {(uint32_t)tg.BX, (uint32_t)tg.BY, (uint32_t)tg.BZ, 1};
  const uint64_t fd[4] = {(uint64_t)W * 3, (uint64_t)H, (uint64_t)out_n0, (uint64_t)B};
  const uint32_t fb[4] = {(uint32_t)TX * 3, (uint32_t)TY, (uint32_t)TZ, 1};
  int rc = encode_f32_4d(&tmv, vol, vd, vb);
  if (rc != NRT_OK) return rc;
  rc = encode_f32_4d(&tmf, flow, fd, fb);
  if (rc != NRT_OK) return rc;
  *used = true;
#define NRT_TILE_CASE(i, tz, ty)                                                                   \
  if (cfg == (i))                                                                                  \
    return method == NRT_LINEAR ? launch_tile<tz, ty, NRT_LINEAR>(tmv, tmf, vol, out, tg, smem, st) \
                                : launch_tile<tz, ty, NRT_NEAREST>(tmv, tmf, vol, out, tg, smem, st);
  NRT_TILE_CASE(0, 8, 16)
  NRT_TILE_CASE(1, 16, 16)
  NRT_TILE_CASE(2, 8, 8)
  NRT_TILE_CASE(3, 4, 8)
#undef NRT_TILE_CASE
  return set_error(NRT_E_ARG, "bad tile config");
}

}  // namespace nrt

using namespace nrt;

extern "C" {

int nrt_interpn_f32(const float* vol, const int32_t* vol_shape, int D, int C, const float* loc,
                    int64_t n_out, int method, int has_fill, float fill, float* out, void* stream) {
  int rc = check_common(D, C, method);
  if (rc != NRT_OK) return rc;
  NRT_REQUIRE(vol && vol_shape && out && (loc || n_out == 0), NRT_E_ARG, "null pointer");
  NRT_REQUIRE(n_out >= 0, NRT_E_ARG, "n_out < 0");
  Geo g;
  int64_t nvox = 1;
  for (int d = 0; d < 3; ++d) {
    g.S[d] = d < D ? vol_shape[d] : 1;
    NRT_REQUIRE(g.S[d] >= 1, NRT_E_ARG, "vol_shape[%d] = %d", d, g.S[d]);
    nvox *= g.S[d];
  }
  NRT_REQUIRE(nvox <= 0x7fffffffLL, NRT_E_SIZE, "volume has %lld voxels (> int32)", (long long)nvox);
  g.src_z0 = 0; g.src_n0 = g.S[0]; g.C = C; g.has_fill = has_fill; g.fill = fill; g.err = nullptr;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define CALL(DD, MM) launch_interpn<DD, MM>(vol, g, loc, n_out, out, st)
  NRT_DISPATCH_D_METHOD(D, method, CALL);
#undef CALL
}

int nrt_warp_f32(const float* vol, const float* flow, float* out, int B, const int32_t* shape, int D,
                 int C, int method, int has_fill, float fill, int src_z0, int src_n0, int out_z0,
                 int out_n0, int halo, int32_t* err_flag, void* stream) {
  int rc = check_common(D, C, method);
  if (rc != NRT_OK) return rc;
  NRT_REQUIRE(vol && flow && out && shape, NRT_E_ARG, "null pointer");
  NRT_REQUIRE(B >= 0, NRT_E_ARG, "B < 0");
  WarpGeo wg;
  int64_t plane = 1;
  for (int d = 0; d < 3; ++d) {
    wg.g.S[d] = d < D ? shape[d] : 1;
    NRT_REQUIRE(wg.g.S[d] >= 1, NRT_E_ARG, "shape[%d] = %d", d, wg.g.S[d]);
    if (d > 0) plane *= wg.g.S[d];
  }
  NRT_REQUIRE(src_z0 >= 0 && src_n0 >= 1 && src_z0 + src_n0 <= shape[0], NRT_E_ARG,
              "resident source planes [%d,%d) outside [0,%d)", src_z0, src_z0 + src_n0, shape[0]);
  NRT_REQUIRE(out_z0 >= 0 && out_n0 >= 0 && out_z0 + out_n0 <= shape[0], NRT_E_ARG,
              "output planes [%d,%d) outside [0,%d)", out_z0, out_z0 + out_n0, shape[0]);
  NRT_REQUIRE(plane * src_n0 <= 0x7fffffffLL && plane * out_n0 * 3 <= 0x7fffffffLL * 4, NRT_E_SIZE,
              "slab too large for int32 indexing");
  if (B == 0 || out_n0 == 0) return NRT_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (D == 3 && C == 1) {
    bool used = false;
    rc = try_tile_path(vol, flow, out, B, shape, method, has_fill, fill, src_z0, src_n0, out_z0, out_n0,
                       halo, err_flag, st, &used);
    if (rc != NRT_OK || used) return rc;
  }
  wg.g.src_z0 = src_z0; wg.g.src_n0 = src_n0; wg.g.C = C;
  wg.g.has_fill = has_fill; wg.g.fill = fill; wg.g.err = err_flag;
  wg.out_z0 = out_z0; wg.out_n0 = out_n0; wg.B = B;
  wg.src_batch_stride = plane * src_n0 * C;
  wg.out_vox = plane * out_n0;
#define CALL(DD, MM) launch_warp_generic<DD, MM>(vol, flow, out, wg, st)
  NRT_DISPATCH_D_METHOD(D, method, CALL);
#undef CALL
}

int nrt_resize_f32(const float* vol, float* out, int B, const int32_t* in_shape, const int32_t* out_shape,
                   int D, int C, int method, int out_z0, int out_n0, void* stream) {
  int rc = check_common(D, C, method);
  if (rc != NRT_OK) return rc;
  NRT_REQUIRE(vol && out && in_shape && out_shape, NRT_E_ARG, "null pointer");
  NRT_REQUIRE(B >= 0, NRT_E_ARG, "B < 0");
  ResizeGeo rg;
  int64_t in_vox = 1, out_plane = 1;
  for (int d = 0; d < 3; ++d) {
    rg.g.S[d] = d < D ? in_shape[d] : 1;
    rg.M[d] = d < D ? out_shape[d] : 1;
    NRT_REQUIRE(rg.g.S[d] >= 1 && rg.M[d] >= 0, NRT_E_ARG, "bad shape at axis %d", d);
    // tf.linspace in fp32: delta = (stop - start) / (num - 1)
    rg.delta[d] = rg.M[d] > 1 ? (float)(rg.g.S[d] - 1) / (float)(rg.M[d] - 1) : 0.0f;
    in_vox *= rg.g.S[d];
    if (d > 0) out_plane *= rg.M[d];
  }
  NRT_REQUIRE(out_z0 >= 0 && out_n0 >= 0 && out_z0 + out_n0 <= rg.M[0], NRT_E_ARG, "output planes out of range");
  NRT_REQUIRE(in_vox <= 0x7fffffffLL && out_plane * out_n0 <= 0x7fffffffLL, NRT_E_SIZE, "volume too large for int32 indexing");
  rg.g.src_z0 = 0; rg.g.src_n0 = rg.g.S[0]; rg.g.C = C; rg.g.has_fill = 0; rg.g.fill = 0.f; rg.g.err = nullptr;
  rg.out_z0 = out_z0; rg.out_n0 = out_n0; rg.B = B;
  rg.src_batch_stride = in_vox * C;
  rg.out_vox = out_plane * out_n0;
  if (B == 0 || rg.out_vox == 0) return NRT_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define CALL(DD, MM) launch_resize<DD, MM>(vol, out, rg, st)
  NRT_DISPATCH_D_METHOD(D, method, CALL);
#undef CALL
}

}  // extern "C"
