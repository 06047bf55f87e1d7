// CUDA-core GEMM: the fp32 verification mode of cseg_gemm (plain TF32 does not hold 1e-4
// through 12-24 layers, so the fp32 mode runs FFMA).  Same epilogue contract as the tcgen05 kernel:
//   C[M,N] = residual + alpha * act(A[M,K] . B[N,K]^T + bias)
// Templated on the operand type so that tests can also feed it bf16 operands as an on-device
// cross-check of the tensor-core kernel.
#include "common.cuh"

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

template <typename TI>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const TI* __restrict__ A, int lda, const TI* __restrict__ B,
                                                        int ldb, int M, int N, int K, const float* __restrict__ bias,
                                                        const void* residual, int ldr, int res_bf16, float alpha,
                                                        int act, int out_bf16, void* C, int ldc) {
  pdl_grid_sync();
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int tx = tid & 15, ty = tid >> 4;  // 16x16 threads, 4x4 outputs each
  float acc[4][4] = {};
  const int lr = tid >> 2, lk = (tid & 3) * 4;  // each thread loads 4 consecutive k of one row
  for (int k0 = 0; k0 < K; k0 += TK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + lk + j;
      const int ra = m0 + lr, rb = n0 + lr;
      As[lk + j][lr] = (ra < M && k < K) ? to_f32(A[(size_t)ra * lda + k]) : 0.f;
      Bs[lk + j][lr] = (rb < N && k < K) ? to_f32(B[(size_t)rb * ldb + k]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col >= N) continue;
      float x = acc[i][j];
      if (bias) x += bias[col];
      x = apply_act(x, act) * alpha;
      if (residual)
        x += res_bf16 ? __bfloat162float(((const bf16*)residual)[(size_t)row * ldr + col])
                      : ((const float*)residual)[(size_t)row * ldr + col];
      if (out_bf16) ((bf16*)C)[(size_t)row * ldc + col] = __float2bfloat16_rn(x);
      else ((float*)C)[(size_t)row * ldc + col] = x;
    }
  }
}

}  // namespace

int cseg_gemm_simt(int in_dtype, const void* A, int lda, const void* B, int ldb, int M, int N, int K,
                   const float* bias, const void* residual, int ldr, int res_dtype, float alpha, int act, int out_dtype,
                   void* C, int ldc, cudaStream_t st) {
  CSEG_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  dim3 grid(cdiv(N, TN), cdiv(M, TM));
  if (in_dtype == CSEG_F32)
    cseg_launch(gemm_simt_kernel<float>, dim3(grid), dim3(256), 0, st, (const float*)A, lda, (const float*)B, ldb, M, N, K, bias, residual,
                                                  ldr, res_dtype == CSEG_BF16, alpha, act, out_dtype == CSEG_BF16, C, ldc);
  else
    cseg_launch(gemm_simt_kernel<bf16>, dim3(grid), dim3(256), 0, st, (const bf16*)A, lda, (const bf16*)B, ldb, M, N, K, bias, residual, ldr,
                                                 res_dtype == CSEG_BF16, alpha, act, out_dtype == CSEG_BF16, C, ldc);
  CSEG_LAUNCH_CHECK("gemm_simt");
  return 0;
}
