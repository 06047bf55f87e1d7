// extern "C" entry points that are not tied to one kernel file (error state, version, GEMM dispatch).
#include "common.cuh"
#include <stdlib.h>

thread_local char g_cseg_err[512] = {0};
std::atomic<long long> g_cseg_launches{0};

bool cseg_pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("CSEG_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

int cseg_gemm_bf16_tc(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const float* bias,
                      const void* residual, int ldr, int res_dtype, float alpha, int act, int out_dtype, void* C, int ldc,
                      cudaStream_t st, int diag_rows = 0);
int cseg_gemm_simt(int in_dtype, const void* A, int lda, const void* B, int ldb, int M, int N, int K,
                   const float* bias, const void* residual, int ldr, int res_dtype, float alpha, int act, int out_dtype, void* C,
                   int ldc, cudaStream_t st);

int cseg_fixup_norm_sim_tc(const void* y, int ldy, const void* W, int ldw, int M, int C, const float* bias, float alpha,
                           const float* text, int Q, const float* cls_bias, int hw, float* logits, cudaStream_t st);

int cseg_basis_logits_tc(const void* s, int lds, int Cb, int n_crops, int hw, int T, int tstride, const void* gram,
                         const void* aux, int ldg, const float* consts, int Q, const float* cls_bias, float* logits,
                         cudaStream_t st);

int cseg_jbu_kernel_fixup_tc(const void* k, int lda, const void* W0, int ldw0, const float* b0, const void* W3, int ldw3,
                             const float* b3, int M, int ldk, void* out, int ldo, cudaStream_t st);

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

int cseg_gemm_blockdiag(const void* A, int lda, const void* B, int ldb, int M, int N, int K, int block_rows, int out_dtype,
                        void* C, int ldc, void* stream) {
  CSEG_REQUIRE(block_rows > 0, "gemm_blockdiag: block_rows=%d", block_rows);
  return cseg_gemm_bf16_tc(A, lda, B, ldb, M, N, K, nullptr, nullptr, 0, CSEG_F32, 1.0f, CSEG_ACT_NONE, out_dtype, C, ldc,
                           (cudaStream_t)stream, block_rows);
}

int cseg_fixup_norm_sim(int dtype, const void* y, int ldy, const void* W, int ldw, int n_crops, int hw, int C,
                        const float* bias, float alpha, const float* text, int Q, const float* cls_logit_bias,
                        float* logits, void* scratch, void* stream) {
  CSEG_REQUIRE(n_crops > 0 && hw > 0 && C > 0 && Q > 0, "fixup_norm_sim: empty problem");
  const long long M = (long long)n_crops * hw;
  CSEG_REQUIRE(M < (1LL << 31), "fixup_norm_sim: too many rows");
  if (dtype == CSEG_BF16) {
    const int rc = cseg_fixup_norm_sim_tc(y, ldy, W, ldw, (int)M, C, bias, alpha, text, Q, cls_logit_bias, hw, logits,
                                          (cudaStream_t)stream);
    if (rc <= 0) return rc;
  }
  // unfused form (fp32 verification mode, or shapes the fused kernel does not cover)
  CSEG_REQUIRE(scratch != nullptr, "fixup_norm_sim: scratch (n_crops*hw*C elements) required for the unfused path");
  int rc = cseg_gemm(dtype, y, ldy, W, ldw, (int)M, C, C, bias, y, ldy, dtype, alpha, CSEG_ACT_NONE, dtype, scratch, C,
                     stream);
  if (rc) return rc;
  return cseg_norm_sim(dtype, scratch, C, n_crops, hw, C, text, Q, cls_logit_bias, logits, stream);
}

int cseg_basis_logits(int dtype, const void* s, int lds, int Cb, int n_crops, int hw, int T, int tstride,
                      const void* gram, const void* aux, int ldg, const float* consts, int Q,
                      const float* cls_logit_bias, float* logits, void* stream) {
  CSEG_REQUIRE(dtype == CSEG_BF16, "basis_logits: bf16 only (the fp32 verification mode upsamples the features directly)");
  CSEG_REQUIRE(s && gram && aux && consts && logits, "basis_logits: null operand");
  return cseg_basis_logits_tc(s, lds, Cb, n_crops, hw, T, tstride, gram, aux, ldg, consts, Q, cls_logit_bias, logits,
                              (cudaStream_t)stream);
}

int cseg_jbu_kernel_fixup(int dtype, const void* k, int lda, const void* W0, int ldw0, const float* b0, const void* W3s,
                          int ldw3, const float* b3s, int M, int ldk, void* out, int ldo, void* stream) {
  CSEG_REQUIRE(dtype == CSEG_BF16, "jbu_kernel_fixup: bf16 only (fp32 mode issues the two GEMMs through cseg_gemm)");
  CSEG_REQUIRE(k && W0 && W3s && out && M > 0, "jbu_kernel_fixup: null operand / empty problem");
  const int rc = cseg_jbu_kernel_fixup_tc(k, lda, W0, ldw0, b0, W3s, ldw3, b3s, M, ldk, out, ldo, (cudaStream_t)stream);
  if (rc == 1) CSEG_FAIL(CSEG_EINVAL, "jbu_kernel_fixup: ldk=%d must be 64 or 128, strides multiples of 8, 16-byte aligned operands", ldk);
  return rc;
}

// test hook: CUDA-core GEMM on bf16 operands (on-device cross-check of the tcgen05 kernel)
int cseg_gemm_reference(int in_dtype, const void* A, int lda, const void* B, int ldb, int M, int N, int K,
                        const float* bias, const void* residual, int ldr, int res_dtype, float alpha, int act, int out_dtype,
                        void* C, int ldc, void* stream) {
  return cseg_gemm_simt(in_dtype, A, lda, B, ldb, M, N, K, bias, residual, ldr, res_dtype, alpha, act, out_dtype, C, ldc,
                        (cudaStream_t)stream);
}

}  // extern "C"
