// Probe (dev tool, GPU box): issue/pipe throughput of scalar FMUL+FADD vs packed FFMA2 on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32x2_probe fp32x2_probe.cu && ./fp32x2_probe
// Each thread runs ITER iterations over 8 independent chains; reports lane-ops per clk per SM.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

constexpr int ITER = 4096;

__device__ __forceinline__ uint64_t pk(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

// mode 0: scalar FMUL + FADD (16 lane-ops per chain-iter pair) ; 1: scalar FFMA ; 2: FFMA2 as mul (c = -0) + FFMA2 as add (b = 1)
// 3: FFMA2 plain
template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, float s, float t) {
  float a[8], b[8];
  uint64_t pa[8], pb[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = threadIdx.x * 1e-3f + i;
    b[i] = 1.0f + i * 1e-3f;
    pa[i] = pk(a[i], a[i] + 0.5f);
    pb[i] = pk(b[i], b[i] + 0.25f);
  }
  const uint64_t one = pk(1.0f, 1.0f), nz = pk(-0.0f, -0.0f), ps = pk(s, s), pt = pk(t, t);
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {
        const float p = __fmul_rn(a[i], s);
        a[i] = __fadd_rn(p, b[i]);
      } else if (MODE == 1) {
        a[i] = fmaf(a[i], s, b[i]);
        b[i] = fmaf(b[i], t, a[i]);
      } else if (MODE == 2) {
        const uint64_t p = fma2(pa[i], ps, nz);
        pa[i] = fma2(p, one, pb[i]);
      } else {
        pa[i] = fma2(pa[i], ps, pb[i]);
        pb[i] = fma2(pb[i], pt, pa[i]);
      }
    }
  }
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(pa[i]));
    float lo2, hi2;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo2), "=f"(hi2) : "l"(pb[i]));
    r += a[i] + b[i] + lo + hi + lo2 + hi2;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, int sms, float* out, double lane_ops_per_inner) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = sms * 8;
  probe<MODE><<<blocks, 256>>>(out, 0.999f, 1.001f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  probe<MODE><<<blocks, 256>>>(out, 0.999f, 1.001f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int clk_khz;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double instr = (double)blocks * 256 * ITER * 8 * 2;          // warp-lane instructions
  const double ops = (double)blocks * 256 * ITER * 8 * lane_ops_per_inner;
  printf("%-34s %8.3f ms  %7.1f lane-instr/clk/SM  %7.1f fp32 results/clk/SM (at %d MHz nominal)\n", name, ms,
         instr / (ms * 1e-3) / (clk_khz * 1e3) / sms, ops / (ms * 1e-3) / (clk_khz * 1e3) / sms, clk_khz / 1000);
}

int main() {
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
  run<0>("scalar FMUL + FADD", sms, out, 2);
  run<1>("scalar FFMA x2", sms, out, 2);
  run<2>("FFMA2 (mul, c=-0) + FFMA2 (add, b=1)", sms, out, 4);
  run<3>("FFMA2 x2", sms, out, 4);
  cudaError_t e = cudaGetLastError();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
