// Probe (dev tool, GPU box): issue rate of the warp-level mma.sync shapes on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe mma_probe.cu && ./mma_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

constexpr int ITER = 2048;

template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out) {
  float d[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
  uint32_t a[4] = {0x3f800000u + threadIdx.x, 0x3f000000u, 0x3e800000u, 0x3f400000u};
  uint32_t b[2] = {0x3f800000u, 0x3f000000u + threadIdx.x};
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
      } else if (MODE == 1) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
      } else if (MODE == 2) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
      } else if (MODE == 3) {
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                     : "r"(a[0]), "r"(a[1]), "r"(b[0]));
      } else {
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                     : "r"(a[0]), "r"(a[1]), "r"(b[0]));
      }
    }
  }
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, int sms, float* out, double fma_per_instr, int blocks_per_sm) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = sms * blocks_per_sm;
  probe<MODE><<<blocks, 256>>>(out);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  probe<MODE><<<blocks, 256>>>(out);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int clk_khz;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double warp_instr = (double)blocks * 8 * ITER * 8;
  const double per_clk_sm = warp_instr / (ms * 1e-3) / (clk_khz * 1e3) / sms;
  printf("%-28s %d CTA/SM %8.3f ms  %6.3f mma/clk/SM  = %7.0f FMA/clk/SM  (%.1f clk per mma per SMSP)\n", name, blocks_per_sm,
         ms, per_clk_sm, per_clk_sm * fma_per_instr, 4.0 / per_clk_sm);
}

int main() {
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
  for (int bps : {2, 8}) {
    run<0>("m16n8k8  tf32", sms, out, 16 * 8 * 8, bps);
    run<3>("m16n8k4  tf32", sms, out, 16 * 8 * 4, bps);
    run<1>("m16n8k16 f16", sms, out, 16 * 8 * 16, bps);
    run<4>("m16n8k8  f16", sms, out, 16 * 8 * 8, bps);
    run<2>("m16n8k16 bf16", sms, out, 16 * 8 * 16, bps);
  }
  printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
