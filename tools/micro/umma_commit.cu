// Does tcgen05.commit after every pair of MMAs slow the issue stream?  (whole warp runs the loop, elect.sync issues)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4); d |= (uint64_t)(lbo >> 4) << 16; d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
template <int MODE>   // 0: commit once at the end; 1: commit after every pair; 2: pair accumulates into two different D
__global__ void __launch_bounds__(128, 1) k(uint32_t idesc, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 96 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  if (threadIdx.x < 32) {
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 48 * 1024;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t st = (uint32_t)(i % 6);
      const uint64_t bd = desc(b0 + (i & 3) * 32, 0, 1024);
      if (elect_one()) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint64_t ad = desc(a0 + st * 8192 + half * 4096, 2048, 1024);   // MN-major A, 16-position strips
          const uint32_t d = tm + (MODE == 2 ? half * 64 : 0) + 0;
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(MODE == 1 ? tm + half * 64 : d), "l"(ad), "l"(bd), "r"(idesc), "r"(i) : "memory");
        }
        if (MODE >= 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
      }
      __syncwarp();
    }
    if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    __syncwarp();
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}
template <int MODE> void run(const char* what) {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 4096;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  long long h = 0;
  for (int rep = 0; rep < 2; ++rep) {
    k<MODE><<<148, 128, 100 * 1024>>>(idesc, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", what, cudaGetErrorString(e)); return; }
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  }
  printf("%-60s %.1f clk per pair of MMAs (M=128, N=64, K=16, MN-major A)\n", what, (double)h / iters);
}
int main() {
  run<0>("two MMAs per iteration, same D, commit at the end:");
  run<2>("two MMAs per iteration, two D, commit every iteration:");
  run<1>("same, (two D) commit every iteration:");
  return 0;
}
