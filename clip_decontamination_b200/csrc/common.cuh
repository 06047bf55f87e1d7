// Shared helpers for the libclipseg kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>

#include "../../include/clipseg.h"

typedef __nv_bfloat16 bf16;

extern thread_local char g_cseg_err[512];
extern std::atomic<long long> g_cseg_launches;

#define CSEG_FAIL(code, ...)                                   \
  do {                                                         \
    snprintf(g_cseg_err, sizeof(g_cseg_err), __VA_ARGS__);     \
    return (code);                                             \
  } while (0)

#define CSEG_REQUIRE(cond, ...)                                \
  do {                                                         \
    if (!(cond)) CSEG_FAIL(CSEG_EINVAL, __VA_ARGS__);          \
  } while (0)

// every launcher ends with this: counts the launch and surfaces launch-time errors
#define CSEG_LAUNCH_CHECK(name)                                                         \
  do {                                                                                  \
    g_cseg_launches.fetch_add(1, std::memory_order_relaxed);                            \
    cudaError_t e_ = cudaGetLastError();                                                \
    if (e_ != cudaSuccess) CSEG_FAIL(CSEG_ECUDA, "%s: %s", name, cudaGetErrorString(e_)); \
  } while (0)

#define CSEG_CUDA(call)                                                                  \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) CSEG_FAIL(CSEG_ECUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

// raise a kernel's dynamic shared-memory limit only when needed (no runtime call on the steady-state
// path, so launch sequences can be captured into CUDA graphs)
#define CSEG_SET_SMEM(kernel, bytes)                                                                      \
  do {                                                                                                    \
    static int cur_[16] = {0};                       /* per device: the attribute is per (function, device) */ \
    int dev_ = 0;                                                                                         \
    cudaGetDevice(&dev_);                                                                                 \
    dev_ &= 15;                                                                                           \
    if ((int)(bytes) > cur_[dev_]) {                                                                      \
      CSEG_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
      cur_[dev_] = (int)(bytes);                                                                          \
    }                                                                                                     \
  } while (0)

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------
// Every kernel of the library is launched with programmatic stream serialisation allowed and starts with
// pdl_grid_sync(): the next kernel of the stream (or of a captured graph) may be scheduled while the tail of the
// previous one drains, runs its input-independent prologue (barrier init, TMEM allocation, tensor-map prefetch),
// and blocks in griddepcontrol.wait until the previous grid has completed and its writes are visible.  Memory
// ordering between consecutive kernels is therefore exactly that of a plain stream; only launch latency and
// prologues overlap.  CSEG_PDL=0 in the environment restores plain launches.
__device__ __forceinline__ void pdl_grid_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
bool cseg_pdl_enabled();
template <typename... KArgs, typename... Args>
inline void cseg_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = cseg_pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // errors surface in CSEG_LAUNCH_CHECK
}

// ---- input image access (cseg_image, include/clipseg.h): normalised RGB value of canvas element (c, Y, x) -----------
struct ImgView {
  const void* data;
  int dtype, H, W, img_h;
  long long stride_img, stride_c, stride_y, stride_x;
  int chan[3];
  float mean[3], std[3];
};
static inline int make_img_view(const cseg_image* d, ImgView& v) {
  if (!d || !d->data || d->H <= 0 || d->W <= 0 || d->img_h <= 0 || d->H % d->img_h != 0) return 1;
  if (d->dtype != CSEG_F32 && d->dtype != CSEG_U8) return 1;
  v.data = d->data; v.dtype = d->dtype; v.H = d->H; v.W = d->W; v.img_h = d->img_h;
  v.stride_img = d->stride_img; v.stride_c = d->stride_c; v.stride_y = d->stride_y; v.stride_x = d->stride_x;
  for (int c = 0; c < 3; ++c) {
    if (d->chan[c] < 0 || d->chan[c] > 2) return 1;
    v.chan[c] = d->chan[c]; v.mean[c] = d->mean[c]; v.std[c] = d->std[c];
    if (d->dtype == CSEG_U8 && !(d->std[c] > 0.f)) return 1;
  }
  return 0;
}
// offset of canvas row Y (image index resolved) without the channel / column terms
__device__ __forceinline__ long long img_row_off(const ImgView& v, int Y) {
  if (v.img_h == v.H) return (long long)Y * v.stride_y;
  const int b = Y / v.img_h;
  return (long long)b * v.stride_img + (long long)(Y - b * v.img_h) * v.stride_y;
}
// per-channel fields are picked with selects on constant indices: indexing the kernel-parameter struct with a runtime channel
// would copy it to local memory (96-184 bytes of stack and a local load per access in the kernels that read the image)
__device__ __forceinline__ int img_chan(const ImgView& v, int c) { return c == 0 ? v.chan[0] : (c == 1 ? v.chan[1] : v.chan[2]); }
__device__ __forceinline__ float img_mean(const ImgView& v, int c) { return c == 0 ? v.mean[0] : (c == 1 ? v.mean[1] : v.mean[2]); }
__device__ __forceinline__ float img_std(const ImgView& v, int c) { return c == 0 ? v.std[0] : (c == 1 ? v.std[1] : v.std[2]); }
__device__ __forceinline__ float img_at(const ImgView& v, int c, long long row_off, int x) {
  const long long off = row_off + (long long)img_chan(v, c) * v.stride_c + (long long)x * v.stride_x;
  if (v.dtype == CSEG_U8)      // SegDataPreProcessor: (x.float() - mean) / std, IEEE division (segmentor.py:64-67)
    return __fdiv_rn((float)reinterpret_cast<const uint8_t*>(v.data)[off] - img_mean(v, c), img_std(v, c));
  return reinterpret_cast<const float*>(v.data)[off];
}

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// exact-erf GELU (nn.GELU default) and QuickGELU (open_clip/transformer.py:35-38)
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
// GELU with erf from Abramowitz-Stegun 7.1.28 (|err| <= 3e-7):  erf(z) = 1 - (1 + a1 z + ... + a6 z^6)^-16.
// One MUFU.RCP and ~16 FMA-pipe instructions; erff() costs a branchy polynomial plus MUFU.EX2 and the
// 7.1.26 form two MUFU ops -- the GELU epilogues of the tensor-core GEMMs are MUFU/issue bound.
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float p = fmaf(0.0000430638f, z, 0.0002765672f);
  p = fmaf(p, z, 0.0001520143f);
  p = fmaf(p, z, 0.0092705272f);
  p = fmaf(p, z, 0.0422820123f);
  p = fmaf(p, z, 0.0705230784f);
  p = fmaf(p, z, 1.0f);
  float r = __frcp_rn(p);
  r *= r; r *= r; r *= r; r *= r;          // p^-16
  const float e = 1.0f - r;                 // erf(|x| / sqrt 2)
  return 0.5f * x * (1.0f + copysignf(e, x));
}
// GELU for bf16 outputs: tanh form on MUFU.TANH, 6 instructions.  |gelu_tanh - gelu_erf| <= 4.8e-4 (at |x| ~ 2.7,
// where half a bf16 ulp is 7.8e-3) and <= 2e-5 for |x| < 0.5, i.e. at least 10x below the rounding of the bf16 store
// that follows; fp32 outputs keep the erf form (gelu_fast / gelu_erf).
__device__ __forceinline__ float gelu_tanh(float x) {
  const float x2 = x * x;
  const float u = x * fmaf(x2, 0.0356774081f, 0.7978845608f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float h = 0.5f * x;
  return fmaf(h, t, h);
}
__device__ __forceinline__ float quick_gelu(float x) { return x / (1.0f + __expf(-1.702f * x)); }
__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == CSEG_ACT_GELU) return gelu_erf(x);
  if (act == CSEG_ACT_QUICKGELU) return quick_gelu(x);
  return x;
}

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline int sm_count() {
  static int n[16] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 15;
  if (n[dev] == 0) {
    cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
    if (n[dev] <= 0) n[dev] = 148;
  }
  return n[dev];
}
