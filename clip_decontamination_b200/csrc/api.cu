// extern "C" entry points that are not tied to one kernel file (error state, version, GEMM dispatch).
#include "common.cuh"

thread_local char g_cseg_err[512] = {0};
std::atomic<long long> g_cseg_launches{0};

int cseg_gemm_bf16_tc(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const float* bias,
                      const void* residual, int ldr, int res_dtype, float alpha, int act, int out_dtype, void* C, int ldc,
                      cudaStream_t st);
int cseg_gemm_simt(int in_dtype, const void* A, int lda, const void* B, int ldb, int M, int N, int K,
                   const float* bias, const void* residual, int ldr, int res_dtype, float alpha, int act, int out_dtype, void* C,
                   int ldc, cudaStream_t st);

extern "C" {

int cseg_version(void) { return CSEG_VERSION; }

int cseg_last_error(char* buf, size_t n) {
  if (!buf || n == 0) return (int)strlen(g_cseg_err);
  strncpy(buf, g_cseg_err, n - 1);
  buf[n - 1] = 0;
  return (int)strlen(buf);
}

long long cseg_launch_count(void) { return g_cseg_launches.load(); }

int cseg_gemm(int in_dtype, const void* A, int lda, const void* B, int ldb, int M, int N, int K, const float* bias,
              const void* residual, int ldr, int res_dtype, float alpha, int act, int out_dtype, void* C, int ldc, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (in_dtype == CSEG_BF16)
    return cseg_gemm_bf16_tc(A, lda, B, ldb, M, N, K, bias, residual, ldr, res_dtype, alpha, act, out_dtype, C, ldc, st);
  if (in_dtype == CSEG_F32)
    return cseg_gemm_simt(CSEG_F32, A, lda, B, ldb, M, N, K, bias, residual, ldr, res_dtype, alpha, act, out_dtype, C, ldc, st);
  CSEG_FAIL(CSEG_EINVAL, "gemm: unknown dtype %d", in_dtype);
}

// test hook: CUDA-core GEMM on bf16 operands (on-device cross-check of the tcgen05 kernel)
int cseg_gemm_reference(int in_dtype, const void* A, int lda, const void* B, int ldb, int M, int N, int K,
                        const float* bias, const void* residual, int ldr, int res_dtype, float alpha, int act, int out_dtype,
                        void* C, int ldc, void* stream) {
  return cseg_gemm_simt(in_dtype, A, lda, B, ldb, M, N, K, bias, residual, ldr, res_dtype, alpha, act, out_dtype, C, ldc,
                        (cudaStream_t)stream);
}

}  // extern "C"
