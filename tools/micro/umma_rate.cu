// tcgen05.mma issue rate by shape / operand major-ness (operands are zeros in shared memory; no loads).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__global__ void __launch_bounds__(128, 1) rate_kernel(uint32_t idesc, int a_mn, int b_mn, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 96 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
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
  if (threadIdx.x == 0) {
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 32 * 1024;
    // K-major: SBO = 1024 (8-row groups), K advance +32 B.  MN-major: LBO = 4096 (64-element chunks), SBO = 1024, K advance 2048 B
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int k = i & 1;
      const uint64_t ad = a_mn ? desc(a0 + k * 2048, 4096, 1024) : desc(a0 + k * 32, 0, 1024);
      const uint64_t bd = b_mn ? desc(b0 + k * 2048, 4096, 1024) : desc(b0 + k * 32, 0, 1024);
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm), "l"(ad), "l"(bd), "r"(idesc), "r"(i)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 4096;
  const int Ns[] = {16, 32, 64, 128, 256};
  for (int n : Ns)
    for (int amn = 0; amn < 2; ++amn)
      for (int bmn = 0; bmn < 2; ++bmn) {
        if (bmn && n % 64) continue;     // MN-major B: whole 64-element chunks
        uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)amn << 15) | ((uint32_t)bmn << 16) |
                         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        long long h = 0;
        for (int rep = 0; rep < 2; ++rep) {
          rate_kernel<<<148, 128, 100 * 1024>>>(idesc, amn, bmn, iters, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("N=%d amn=%d bmn=%d: %s\n", n, amn, bmn, cudaGetErrorString(e)); return 1; }
          cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        }
        printf("M=128 N=%3d A %s B %s: %.1f clk/MMA (floor %d)\n", n, amn ? "MN" : "K ", bmn ? "MN" : "K ", (double)h / iters, n / 2);
      }
  return 0;
}
