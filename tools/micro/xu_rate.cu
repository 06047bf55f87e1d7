// Throughput of the XU-pipe instruction classes the softmax / conversion-heavy kernels depend on (B200, sm_100a):
// MUFU.EX2, F2FP.BF16.F32.PACK_AB (two floats -> bf16x2), F2F-style single conversions, and, for comparison, FFMA.
// Every warp of a full SM issues independent instructions of one class; reports lanes per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xu_rate xu_rate.cu && ./xu_rate
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int ITER = 4096, UNROLL = 8;

template <int KIND>
__global__ void __launch_bounds__(1024) rate_kernel(float* out, long long* clk, float seed) {
  float v[UNROLL];
  unsigned acc = 0;
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) v[u] = seed + 0.001f * (threadIdx.x + u);
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < ITER; ++i) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (KIND == 0) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[u]));
      } else if (KIND == 1) {
        unsigned r;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(v[u]), "f"(v[(u + 1) % UNROLL]));
        acc ^= r;
      } else if (KIND == 2) {
        unsigned short r;
        asm volatile("cvt.rn.bf16.f32 %0, %1;" : "=h"(r) : "f"(v[u]));
        acc ^= r;
      } else if (KIND == 3) {
        unsigned r;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(v[u]), "f"(v[(u + 1) % UNROLL]));
        acc ^= r;
      } else {
        asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(v[u]));
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) s += v[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)acc;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(const char* name) {
  float* out;
  long long* clk;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&clk, 148 * sizeof(long long));
  rate_kernel<KIND><<<148, 1024>>>(out, clk, 0.5f);
  rate_kernel<KIND><<<148, 1024>>>(out, clk, 0.5f);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0;
  for (int i = 0; i < 148; ++i) c += h[i];
  c /= 148;
  const double lanes = 1024.0 * ITER * UNROLL;
  printf("%-34s %8.0f clk  -> %6.2f lanes/clk/SM  (%5.2f clk per warp instruction per SMSP)\n", name, c, lanes / c, c / (ITER * UNROLL * 8.0));
  cudaFree(out);
  cudaFree(clk);
}

int main() {
  run<0>("MUFU.EX2 (ex2.approx.ftz.f32)");
  run<1>("F2FP.BF16.PACK_AB (cvt bf16x2)");
  run<2>("cvt.rn.bf16.f32 (single)");
  run<3>("F2FP.F16.PACK_AB (cvt f16x2)");
  run<4>("FFMA");
  return 0;
}
