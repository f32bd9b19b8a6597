// Dev probe (GPU box): isolates the 4-D TMA tensor load used by warp3d_tile_kernel.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_probe tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../../neurite_b200/csrc/nrt_common.cuh"
using namespace nrt;
namespace nrt { int set_error(int s, const char*, ...) { return s; } }

__global__ void probe(const __grid_constant__ CUtensorMap tm, float* out, int n, int c0, int c1, int c2, int c3) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* s = reinterpret_cast<float*>(smem);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + ((n * 4 + 127) / 128) * 128);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
    mbar_expect_tx(bar, n * 4);
    tma_load_4d(s, &tm, bar, c0, c1, c2, c3);
  }
  __syncthreads();
  mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = s[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int run(EncodeTiledFn enc, int W, int H, int D, int B, int bx, int by, int bz, int c0, int c1, int c2) {
  size_t n = (size_t)W * H * D * B;
  std::vector<float> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = (float)i;
  float *d, *o;
  cudaMalloc(&d, n * 4);
  cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
  int nb = bx * by * bz;
  cudaMalloc(&o, nb * 4);
  CUtensorMap tm;
  cuuint64_t gd[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B};
  cuuint64_t gs[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * D * 4};
  cuuint32_t bb[4] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz, 1}, es[4] = {1, 1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, gd, gs, bb, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("case W%d H%d D%d box %dx%dx%d at (%d,%d,%d): encode=%d ", W, H, D, bx, by, bz, c0, c1, c2, (int)r);
  size_t smem = ((nb * 4 + 127) / 128) * 128 + 16;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<<<1, 128, smem>>>(tm, o, nb, c0, c1, c2, 0);
  cudaError_t e = cudaDeviceSynchronize();
  printf("sync=%s ", cudaGetErrorString(e));
  if (e != cudaSuccess) { printf("\n"); return 1; }
  std::vector<float> ho(nb);
  cudaMemcpy(ho.data(), o, nb * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int z = 0; z < bz; ++z) for (int y = 0; y < by; ++y) for (int x = 0; x < bx; ++x) {
    int gx = c0 + x, gy = c1 + y, gz = c2 + z;
    float want = (gx < 0 || gx >= W || gy < 0 || gy >= H || gz < 0 || gz >= D) ? 0.f : h[((size_t)gz * H + gy) * W + gx];
    if (ho[(z * by + y) * bx + x] != want) ++bad;
  }
  printf("mismatches=%d\n", bad);
  cudaFree(d); cudaFree(o);
  return bad != 0;
}

int main(int argc, char** argv) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  if (argc < 10) { printf("usage: W H D bx by bz c0 c1 c2\n"); return 2; }
  int a[9];
  for (int i = 0; i < 9; ++i) a[i] = atoi(argv[i + 1]);
  return run(enc, a[0], a[1], a[2], 1, a[3], a[4], a[5], a[6], a[7], a[8]);
}
