// Microbenchmark: peak rate of legacy mma.sync m16n8k16 bf16 on sm_100a (registers only).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, int iters, int nacc) {
  float d[16][4];
  for (int i = 0; i < 16; ++i) for (int e = 0; e < 4; ++e) d[i][e] = 0.f;
  unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0; for (int i = 0; i < 16; ++i) for (int e = 0; e < 4; ++e) s += d[i][e];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
  for (int warps = 4; warps <= 32; warps *= 2) {
    for (int ctas = 1; ctas <= 2; ++ctas) {
      int iters = 20000;
      k<<<148 * ctas, warps * 32>>>(out, 100, 16); cudaDeviceSynchronize();
      cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
      cudaEventRecord(s); k<<<148 * ctas, warps * 32>>>(out, iters, 16); cudaEventRecord(e); cudaEventSynchronize(e);
      float ms; cudaEventElapsedTime(&ms, s, e);
      double flops = 2.0 * 16 * 8 * 16 * 16.0 * iters * warps * 148 * ctas;
      printf("warps/CTA %2d CTAs/SM %d: %.3f ms  %.1f TFLOP/s  (%.0f MAC/clk/SM @1.9GHz)\n", warps, ctas, ms, flops / ms / 1e9,
             flops / 2 / (ms * 1e-3) / 148 / 1.9e9);
    }
  }
  return 0;
}
