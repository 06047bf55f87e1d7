// SimFeatUp joint-bilateral upsampler (simfeatup_dev/upsamplers.py:202-325), channel-last layout.
//   guidance  : adaptive_avg_pool2d of the crop                           (:316)
//   range_proj: conv1x1(3->32) . GELU . conv1x1(32->32)                    (:209-214)
//   range_kernel: (2r+1)^2 reflect-padded key.query logits -> softmax * Gaussian, renormalised
//                                                                          (:230-251,258-262)
//   [fixup_proj: two 1x1 convs = two cseg_gemm calls with GELU / residual epilogues, :264]
//   apply     : bicubic x2 + reflect pad + adaptive conv                   (:268-274, :14-25)
#include "common.cuh"
#include "jbu_share.cuh"
#include <stdlib.h>

namespace {

__device__ __forceinline__ int reflect_idx(int i, int n) {  // F.pad(mode='reflect'), single bounce
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// ---------------------------------------------------------------------------------------------
__global__ void guidance_kernel(const ImgView img, const int32_t* __restrict__ wins,
                                int n_crops, int crop_h, int crop_w, int pad_top, int pad_left, int gh, int gw,
                                float4* __restrict__ guid) {
  pdl_grid_sync();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_crops * gh * gw) return;
  const int gx = idx % gw, gy = (idx / gw) % gh, crop = idx / (gw * gh);
  const int y1 = wins[crop * 4], x1 = wins[crop * 4 + 1], wh = wins[crop * 4 + 2], ww = wins[crop * 4 + 3];
  // adaptive pooling window: [floor(i*in/out), ceil((i+1)*in/out))
  const int ys = (gy * crop_h) / gh, ye = ((gy + 1) * crop_h + gh - 1) / gh;
  const int xs = (gx * crop_w) / gw, xe = ((gx + 1) * crop_w + gw - 1) / gw;
  float acc[3] = {0.f, 0.f, 0.f};
  for (int c = 0; c < 3; ++c)
    for (int y = ys; y < ye; ++y)
      for (int x = xs; x < xe; ++x) {
        const int cy = y - pad_top, cx = x - pad_left;
        if (cy >= 0 && cy < wh && cx >= 0 && cx < ww) acc[c] += img_at(img, c, img_row_off(img, y1 + cy), x1 + cx);
      }
  const float inv = 1.0f / (float)((ye - ys) * (xe - xs));
  guid[idx] = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, 0.f);
}

// ---------------------------------------------------------------------------------------------
template <int KD>
__global__ void __launch_bounds__(256) range_proj_kernel(const float4* __restrict__ guid, int n_pix,
                                                         const float* __restrict__ w0, const float* __restrict__ b0,
                                                         const float* __restrict__ w3, const float* __restrict__ b3,
                                                         float* __restrict__ proj) {
  pdl_grid_sync();
  __shared__ float sw0[KD * 3], sb0[KD], sw3[KD * KD], sb3[KD];
  for (int i = threadIdx.x; i < KD * 3; i += blockDim.x) sw0[i] = w0[i];
  for (int i = threadIdx.x; i < KD * KD; i += blockDim.x) sw3[i] = w3[i];
  for (int i = threadIdx.x; i < KD; i += blockDim.x) { sb0[i] = b0[i]; sb3[i] = b3[i]; }
  __syncthreads();
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= n_pix) return;
  const float4 g = guid[pix];
  float h[KD];
#pragma unroll
  for (int k = 0; k < KD; ++k) h[k] = gelu_erf(sw0[k * 3] * g.x + sw0[k * 3 + 1] * g.y + sw0[k * 3 + 2] * g.z + sb0[k]);
  float* o = proj + (size_t)pix * KD;
#pragma unroll 4
  for (int k = 0; k < KD; ++k) {
    float a = sb3[k];
#pragma unroll
    for (int j = 0; j < KD; ++j) a = fmaf(sw3[k * KD + j], h[j], a);
    o[k] = a;
  }
}

// ---------------------------------------------------------------------------------------------
// One thread per pixel; a 16x16 pixel tile's (16+2R)^2 halo of projections is staged in shared memory
// as channel planes (conflict-free for x-adjacent lanes).
// ---------------------------------------------------------------------------------------------
template <typename T, int R, int KD>
__global__ void __launch_bounds__(256) range_kernel_kernel(const float* __restrict__ proj,
                                                           const float4* __restrict__ guid, int gh, int gw,
                                                           float pos_temp, float inv2s2, T* __restrict__ kern,
                                                           int kwidth, int ldk) {
  pdl_grid_sync();
  constexpr int DIA = 2 * R + 1, D2 = DIA * DIA, TS = 16, HS = TS + 2 * R;
  extern __shared__ float sp[];  // [KD][HS*HS]
  const int crop = blockIdx.z, ty0 = blockIdx.y * TS, tx0 = blockIdx.x * TS;
  const float* pc = proj + (size_t)crop * gh * gw * KD;
  for (int e = threadIdx.x; e < HS * HS * KD; e += blockDim.x) {
    const int k = e % KD, pos = e / KD;
    const int hy = pos / HS, hx = pos % HS;
    const int y = reflect_idx(min(ty0 + hy - R, gh - 1 + R), gh), x = reflect_idx(min(tx0 + hx - R, gw - 1 + R), gw);
    sp[k * (HS * HS) + pos] = pc[((size_t)y * gw + x) * KD + k];
  }
  __syncthreads();
  const int lx = threadIdx.x % TS, ly = threadIdx.x / TS;
  const int y = ty0 + ly, x = tx0 + lx;
  if (y >= gh || x >= gw) return;
  float q[KD];
#pragma unroll
  for (int k = 0; k < KD; ++k) q[k] = sp[k * (HS * HS) + (ly + R) * HS + lx + R];
  float lg[D2];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < DIA; ++i)
#pragma unroll
    for (int j = 0; j < DIA; ++j) {
      float a = 0.f;
      const int pos = (ly + i) * HS + lx + j;
#pragma unroll
      for (int k = 0; k < KD; ++k) a = fmaf(sp[k * (HS * HS) + pos], q[k], a);
      a *= pos_temp;
      lg[i * DIA + j] = a;
      m = fmaxf(m, a);
    }
  float se = 0.f;
#pragma unroll
  for (int t = 0; t < D2; ++t) {
    lg[t] = __expf(lg[t] - m);
    se += lg[t];
  }
  const float inv_se = 1.0f / se;
  float sc = 0.f;
#pragma unroll
  for (int i = 0; i < DIA; ++i)
#pragma unroll
    for (int j = 0; j < DIA; ++j) {
      // get_spatial_kernel: coordinates linspace(-1, 1, DIA) on both axes
      const float dy = -1.f + 2.f * i / (DIA - 1), dx = -1.f + 2.f * j / (DIA - 1);
      const float g = __expf(-(dy * dy + dx * dx) * inv2s2);
      lg[i * DIA + j] = lg[i * DIA + j] * inv_se * g;
      sc += lg[i * DIA + j];
    }
  const float inv_sc = 1.0f / fmaxf(sc, 1e-7f);
  const size_t pix = ((size_t)crop * gh + y) * gw + x;
  T* o = kern + pix * ldk;
#pragma unroll
  for (int t = 0; t < D2; ++t) o[t] = from_f32<T>(lg[t] * inv_sc);
  const float4 g = guid[pix];
  o[D2] = from_f32<T>(g.x);
  o[D2 + 1] = from_f32<T>(g.y);
  o[D2 + 2] = from_f32<T>(g.z);
  for (int t = D2 + 3; t < kwidth; ++t) o[t] = from_f32<T>(0.f);
}

// ---------------------------------------------------------------------------------------------
// bicubic x2 upsample, align_corners=False, A=-0.75 (torch upsample_bicubic2d), channel-last
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cubic_coeffs(float t, float (&c)[4]) {
  const float A = -0.75f;
  const float x0 = t + 1.f, x1 = t, x2 = 1.f - t, x3 = 2.f - t;
  c[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  c[1] = ((A + 2.f) * x1 - (A + 3.f)) * x1 * x1 + 1.f;
  c[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  c[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}

// 8-channel vector load/store helpers (16 B for bf16, 2 x 16 B for fp32)
template <typename T> struct V8;
template <> struct V8<bf16> {
  static __device__ __forceinline__ void ld(const bf16* p, float (&v)[8]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void st(bf16* p, const float (&v)[8]) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};
template <> struct V8<float> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void st(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};

// One thread = a 2x2 block of source pixels x 4 channels -> the 4x4 output block (4y0..4y0+3, 4x0..4x0+3).
// Output row 2y samples source rows y-2..y+1 with t = 0.75, row 2y+1 samples y-1..y+2 with t = 0.25
// (src = (dst + 0.5) / 2 - 0.5), so the block needs a clamped 6x6 neighbourhood: 36 loads for 16 outputs
// (2.25 per output; the one-output-per-thread form needs 16).  Horizontal pass first, then vertical
// (torch order).  4 channels per thread keep the 16 x 4 accumulators in registers.
template <typename T> struct V4;
template <> struct V4<bf16> {
  static __device__ __forceinline__ void ld(const bf16* p, float (&v)[4]) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
    const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
  static __device__ __forceinline__ void st(bf16* p, const float (&v)[4]) {
    uint2 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
    h[0] = __floats2bfloat162_rn(v[0], v[1]);
    h[1] = __floats2bfloat162_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = u;
  }
};
template <> struct V4<float> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  }
  static __device__ __forceinline__ void st(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};

template <typename T>
__global__ void __launch_bounds__(256) bicubic2x_kernel(const T* __restrict__ src, int n_crops, int h, int w, int C,
                                                        T* __restrict__ dst) {
  pdl_grid_sync();
  const int cg = C / 4, hb = (h + 1) / 2, wb = (w + 1) / 2;
  const long long total = (long long)n_crops * hb * wb * cg;
  float cA[4], cB[4];
  cubic_coeffs(0.75f, cA);   // even output index
  cubic_coeffs(0.25f, cB);   // odd output index
  const int W2 = 2 * w, H2 = 2 * h;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % cg);
    long long r = idx / cg;
    const int bx = (int)(r % wb);
    r /= wb;
    const int by = (int)(r % hb), crop = (int)(r / hb);
    const int y0 = 2 * by, x0 = 2 * bx;                 // source block origin
    const T* sb = src + (size_t)crop * h * w * C + g * 4;
    float o[4][4][4];                                   // [out row][out col][channel]
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[a][b][e] = 0.f;
#pragma unroll
    for (int dy = -2; dy <= 3; ++dy) {                  // source rows y0-2 .. y0+3
      const int yy = min(max(y0 + dy, 0), h - 1);
      float hz[4][4];                                   // horizontally interpolated, 4 output columns
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int e = 0; e < 4; ++e) hz[b][e] = 0.f;
#pragma unroll
      for (int dx = -2; dx <= 3; ++dx) {                // source cols x0-2 .. x0+3
        const int xx = min(max(x0 + dx, 0), w - 1);
        float v[4];
        V4<T>::ld(sb + ((size_t)yy * w + xx) * C, v);
        // output col 0 (= 2*x0):   src x0-2..x0+1 (cA);  col 1 (2*x0+1): x0-1..x0+2 (cB)
        // output col 2 (= 2*x0+2): src x0-1..x0+2 (cA);  col 3:           x0..x0+3   (cB)
        if (dx <= 1) {
#pragma unroll
          for (int e = 0; e < 4; ++e) hz[0][e] = fmaf(cA[dx + 2], v[e], hz[0][e]);
        }
        if (dx >= -1 && dx <= 2) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            hz[1][e] = fmaf(cB[dx + 1], v[e], hz[1][e]);
            hz[2][e] = fmaf(cA[dx + 1], v[e], hz[2][e]);
          }
        }
        if (dx >= 0) {
#pragma unroll
          for (int e = 0; e < 4; ++e) hz[3][e] = fmaf(cB[dx], v[e], hz[3][e]);
        }
      }
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (dy <= 1) o[0][b][e] = fmaf(cA[dy + 2], hz[b][e], o[0][b][e]);
          if (dy >= -1 && dy <= 2) {
            o[1][b][e] = fmaf(cB[dy + 1], hz[b][e], o[1][b][e]);
            o[2][b][e] = fmaf(cA[dy + 1], hz[b][e], o[2][b][e]);
          }
          if (dy >= 0) o[3][b][e] = fmaf(cB[dy], hz[b][e], o[3][b][e]);
        }
    }
    T* ob = dst + (size_t)crop * H2 * W2 * C + g * 4;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int Y = 2 * y0 + a, X = 2 * x0 + b;
        if (Y < H2 && X < W2) V4<T>::st(ob + ((size_t)Y * W2 + X) * C, o[a][b]);
      }
  }
}

// ---------------------------------------------------------------------------------------------
// adaptive conv over the (virtually) reflect-padded high-res source:
//   out[p, c] = sum_t hr[reflect(p + t)][c] * kern[p][t]
// One thread per (pixel, 8-channel group); threads of a warp share the pixel's weights through L1.
// ---------------------------------------------------------------------------------------------
template <typename T, int R>
__global__ void __launch_bounds__(256) adaptive_conv_kernel(const T* __restrict__ hr, int n_crops, int H2, int W2, int C,
                                                            const T* __restrict__ kern, int ldk, T* __restrict__ dst) {
  pdl_grid_sync();
  constexpr int DIA = 2 * R + 1;
  const int cg = C / 8;
  const long long total = (long long)n_crops * H2 * W2 * cg;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % cg);
    const long long pix = idx / cg;
    const int X = (int)(pix % W2), Y = (int)((pix / W2) % H2), crop = (int)(pix / ((long long)W2 * H2));
    const T* kp = kern + pix * ldk;
    const T* hb = hr + (size_t)crop * H2 * W2 * C + g * 8;
    float acc[8] = {};
#pragma unroll 1
    for (int i = 0; i < DIA; ++i) {
      const int yy = reflect_idx(Y + i - R, H2);
#pragma unroll
      for (int j = 0; j < DIA; ++j) {
        const int xx = reflect_idx(X + j - R, W2);
        const float wv = to_f32(kp[i * DIA + j]);
        float v[8];
        V8<T>::ld(hb + ((size_t)yy * W2 + xx) * C, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(v[e], wv, acc[e]);
      }
    }
    V8<T>::st(dst + pix * C + g * 8, acc);
  }
}

}  // namespace

int cseg_jbu_guidance_proj_f16(const ImgView& img, const int32_t* windows, int n_crops, int crop_h, int crop_w,
                               int pad_top, int pad_left, int gh, int gw, const float* w0, const float* b0, const float* w3,
                               const float* b3, float* guid, void* proj, cudaStream_t st);
int cseg_jbu_range_proj_f16(const float* guid, int n_pix, const float* w0, const float* b0, const float* w3,
                            const float* b3, void* proj, cudaStream_t st);
int cseg_jbu_range_kernel_mma(const void* proj_f16, const float* guid, int n_crops, int gh, int gw, int radius,
                              float pos_temp, float inv2s2, void* kern, int kwidth, int ldk, cudaStream_t st,
                              const ShareGeom* sg = nullptr, int fb = 0);

extern "C" {

int cseg_jbu_guidance(const cseg_image* img_desc, const int32_t* windows, int n_crops, int crop_h, int crop_w,
                      int pad_top, int pad_left, int gh, int gw, float* guid, void* stream) {
  ImgView img;
  CSEG_REQUIRE(make_img_view(img_desc, img) == 0, "jbu_guidance: bad image descriptor");
  CSEG_REQUIRE(n_crops > 0 && gh > 0 && gw > 0 && gh <= crop_h && gw <= crop_w, "jbu_guidance: bad shape");
  const int n = n_crops * gh * gw;
  cseg_launch(guidance_kernel, dim3(cdiv(n, 256)), dim3(256), 0, (cudaStream_t)stream, img, windows, n_crops, crop_h, crop_w, pad_top,
                                                                  pad_left, gh, gw, (float4*)guid);
  CSEG_LAUNCH_CHECK("jbu_guidance");
  return 0;
}

int cseg_jbu_guidance_proj(const cseg_image* img_desc, const int32_t* windows, int n_crops, int crop_h, int crop_w,
                           int pad_top, int pad_left, int gh, int gw, int key_dim, const float* w0, const float* b0,
                           const float* w3, const float* b3, float* guid, int proj_dtype, void* proj, void* stream) {
  ImgView img;
  CSEG_REQUIRE(make_img_view(img_desc, img) == 0, "jbu_guidance_proj: bad image descriptor");
  CSEG_REQUIRE(n_crops > 0 && gh > 0 && gw > 0 && gh <= crop_h && gw <= crop_w, "jbu_guidance_proj: bad shape");
  CSEG_REQUIRE(key_dim == 32 && proj_dtype == CSEG_F16, "jbu_guidance_proj: key_dim 32 / fp16 projections only (bf16 pipeline)");
  CSEG_REQUIRE(gw >= 16, "jbu_guidance_proj: gw=%d must be >= 16 (call cseg_jbu_guidance + cseg_jbu_range_proj instead)", gw);
  return cseg_jbu_guidance_proj_f16(img, windows, n_crops, crop_h, crop_w, pad_top, pad_left, gh, gw, w0, b0, w3, b3, guid,
                                    proj, (cudaStream_t)stream);
}

int cseg_jbu_range_proj(const float* guid, int n_pix, int key_dim, const float* w0, const float* b0, const float* w3,
                        const float* b3, int proj_dtype, void* proj, void* stream) {
  CSEG_REQUIRE(n_pix > 0, "jbu_range_proj: empty");
  CSEG_REQUIRE(key_dim == 32, "jbu_range_proj: key_dim=%d (only 32, simfeatup_dev/upsamplers.py:282-308)", key_dim);
  if (proj_dtype == CSEG_F16) return cseg_jbu_range_proj_f16(guid, n_pix, w0, b0, w3, b3, proj, (cudaStream_t)stream);
  CSEG_REQUIRE(proj_dtype == CSEG_F32, "jbu_range_proj: proj_dtype must be CSEG_F32 or CSEG_F16");
  cseg_launch(range_proj_kernel<32>, dim3(cdiv(n_pix, 256)), dim3(256), 0, (cudaStream_t)stream, (const float4*)guid, n_pix, w0, b0, w3, b3,
                                                                            (float*)proj);
  CSEG_LAUNCH_CHECK("jbu_range_proj");
  return 0;
}

}  // extern "C"

template <typename T, int R>
static int launch_range_kernel(const float* proj, const float* guid, int n_crops, int gh, int gw, float pos_temp,
                               float inv2s2, void* kern, int kwidth, int ldk, cudaStream_t st) {
  constexpr int HS = 16 + 2 * R;
  const size_t smem = (size_t)32 * HS * HS * sizeof(float);
  CSEG_SET_SMEM((range_kernel_kernel<T, R, 32>), smem);
  dim3 grid(cdiv(gw, 16), cdiv(gh, 16), n_crops);
  cseg_launch(range_kernel_kernel<T, R, 32>, dim3(grid), dim3(256), smem, st, proj, (const float4*)guid, gh, gw, pos_temp, inv2s2, (T*)kern,
                                                         kwidth, ldk);
  CSEG_LAUNCH_CHECK("jbu_range_kernel");
  return 0;
}

int cseg_jbu_adaptive_conv_tc(const bf16* hr, int n_crops, int H2, int W2, int C, const bf16* kern, int ldk, int radius,
                              bf16* dst, cudaStream_t st);
int cseg_jbu_apply_fused(const bf16* src, int n_crops, int h, int w, int C, const bf16* kern, int ldk, int radius,
                         bf16* dst, void* scratch, cudaStream_t st);
int cseg_jbu_composite_image_tc(const bf16* kern_img, int ldk, int ih, int iw, int gh, int gw, int radius, bf16* kc_img,
                                void* tabs, cudaStream_t st);
int cseg_jbu_apply_shared_tc(const bf16* src, int n_crops, int h, int w, int C, const bf16* kern_border, const bf16* kern_img,
                             const bf16* kc_img, const ShareGeom& sg, int ldk, int radius, bf16* dst, void* scratch,
                             cudaStream_t st);
static bool apply_fused_enabled() {  // CSEG_APPLY_FUSED=0 selects bicubic2x + the stand-alone conv (A/B measurements)
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("CSEG_APPLY_FUSED");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}
static bool conv_tc_enabled() {      // CSEG_CONV_TC=0 selects the mma.sync kernel (A/B measurements)
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("CSEG_CONV_TC");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}
int cseg_jbu_adaptive_conv_mma(const bf16* hr, int n_crops, int H2, int W2, int C, const bf16* kern, int ldk,
                               int radius, bf16* dst, cudaStream_t st);

template <typename T>
static int launch_apply(const void* src, int n_crops, int h, int w, int C, const void* kern, int ldk, int radius,
                        void* dst, void* hr_scratch, cudaStream_t st) {
  const int H2 = 2 * h, W2 = 2 * w;
  if (sizeof(T) == 2 && apply_fused_enabled()) {   // bicubic folded into the kernel weights: one tcgen05 kernel on the low-res source
    const int rc = cseg_jbu_apply_fused((const bf16*)src, n_crops, h, w, C, (const bf16*)kern, ldk, radius, (bf16*)dst,
                                        hr_scratch, st);
    if (rc <= 0) return rc;
  }
  const long long tot_b = (long long)n_crops * ((h + 1) / 2) * ((w + 1) / 2) * (C / 4);
  cseg_launch(bicubic2x_kernel<T>, dim3((int)std::min<long long>((tot_b + 255) / 256, (long long)sm_count() * 32)), dim3(256), 0, st, 
      (const T*)src, n_crops, h, w, C, (T*)hr_scratch);
  CSEG_LAUNCH_CHECK("jbu_bicubic2x");
  if (sizeof(T) == 2 && conv_tc_enabled()) {                                      // tcgen05 banded GEMM (C % 128 == 0)
    const int rc = cseg_jbu_adaptive_conv_tc((const bf16*)hr_scratch, n_crops, H2, W2, C, (const bf16*)kern, ldk, radius,
                                             (bf16*)dst, st);
    if (rc <= 0) return rc;
  }
  if (sizeof(T) == 2 && C % 64 == 0 && C >= 128 && ldk % 8 == 0 && ldk <= 128)   // mma.sync banded GEMM path
    return cseg_jbu_adaptive_conv_mma((const bf16*)hr_scratch, n_crops, H2, W2, C, (const bf16*)kern, ldk, radius,
                                      (bf16*)dst, st);
  const long long tot = (long long)n_crops * H2 * W2 * (C / 8);
  const int blocks = (int)std::min<long long>((tot + 255) / 256, (long long)sm_count() * 64);
  if (radius == 5)
    cseg_launch(adaptive_conv_kernel<T, 5>, dim3(blocks), dim3(256), 0, st, (const T*)hr_scratch, n_crops, H2, W2, C, (const T*)kern, ldk,
                                                       (T*)dst);
  else
    cseg_launch(adaptive_conv_kernel<T, 3>, dim3(blocks), dim3(256), 0, st, (const T*)hr_scratch, n_crops, H2, W2, C, (const T*)kern, ldk,
                                                       (T*)dst);
  CSEG_LAUNCH_CHECK("jbu_adaptive_conv");
  return 0;
}

extern "C" {

int cseg_jbu_range_kernel(int proj_dtype, const void* proj_v, const float* guid, int n_crops, int gh, int gw, int key_dim,
                          int radius, float range_temp, float sigma_spatial, int out_dtype, void* kern, int kwidth,
                          int ldk, void* stream) {
  CSEG_REQUIRE(n_crops > 0 && gh > radius && gw > radius, "jbu_range_kernel: grid %dx%d too small for radius %d", gh, gw, radius);
  CSEG_REQUIRE(key_dim == 32, "jbu_range_kernel: key_dim=%d (only 32)", key_dim);
  CSEG_REQUIRE(radius == 3 || radius == 5, "jbu_range_kernel: radius=%d (3 = jbu_stack, 5 = jbu_one)", radius);
  const int d2 = (2 * radius + 1) * (2 * radius + 1);
  CSEG_REQUIRE(kwidth >= d2 + 3 && ldk >= kwidth, "jbu_range_kernel: kwidth=%d (>= %d), ldk=%d (>= kwidth)", kwidth, d2 + 3, ldk);
  // pos_temp = exp(range_temp).clamp(1e-4, 1e4)   (simfeatup_dev/upsamplers.py:237)
  const float pos_temp = fminf(fmaxf(expf(range_temp), 1e-4f), 1e4f);
  const float inv2s2 = 1.0f / (2.0f * sigma_spatial * sigma_spatial);
  cudaStream_t st = (cudaStream_t)stream;
  if (proj_dtype == CSEG_F16) {   // tensor-core path
    CSEG_REQUIRE(out_dtype == CSEG_BF16, "jbu_range_kernel: fp16 projections go with bf16 kernels");
    const int rc = cseg_jbu_range_kernel_mma(proj_v, guid, n_crops, gh, gw, radius, pos_temp, inv2s2, kern, kwidth, ldk, st);
    if (rc == 1) CSEG_FAIL(CSEG_EUNSUPPORTED, "jbu_range_kernel(f16): radius=%d kwidth=%d not covered (5/128, 3/64)", radius, kwidth);
    return rc;
  }
  CSEG_REQUIRE(proj_dtype == CSEG_F32, "jbu_range_kernel: proj_dtype must be CSEG_F32 or CSEG_F16");
  const float* proj = (const float*)proj_v;
  if (out_dtype == CSEG_BF16) {
    if (radius == 5) return launch_range_kernel<bf16, 5>(proj, guid, n_crops, gh, gw, pos_temp, inv2s2, kern, kwidth, ldk, st);
    return launch_range_kernel<bf16, 3>(proj, guid, n_crops, gh, gw, pos_temp, inv2s2, kern, kwidth, ldk, st);
  }
  if (radius == 5) return launch_range_kernel<float, 5>(proj, guid, n_crops, gh, gw, pos_temp, inv2s2, kern, kwidth, ldk, st);
  return launch_range_kernel<float, 3>(proj, guid, n_crops, gh, gw, pos_temp, inv2s2, kern, kwidth, ldk, st);
}

static int share_geom(const cseg_jbu_share* sh, ShareGeom& sg, const char* who) {
  CSEG_REQUIRE(sh && sh->windows && sh->shift >= 0 && sh->shift <= 4 && sh->pitch > 0, "%s: bad cseg_jbu_share", who);
  sg.wins = sh->windows;
  sg.shift = sh->shift;
  sg.pitch = sh->pitch;
  return 0;
}

int cseg_jbu_share_rows(int gh, int gw, int fb) { return border_rows(gh, gw, fb); }

int cseg_jbu_range_kernel_border(const void* proj_img, const float* guid_img, const cseg_jbu_share* share, int n_crops,
                                 int gh, int gw, int radius, float range_temp, float sigma_spatial, void* kern_border,
                                 int kwidth, int ldk, void* stream) {
  ShareGeom sg;
  if (int rc = share_geom(share, sg, "jbu_range_kernel_border")) return rc;
  CSEG_REQUIRE(n_crops > 0 && gh > 2 * CSEG_JBU_FB_RANGE && gw >= 32, "jbu_range_kernel_border: region %dx%d too small", gh, gw);
  const float pos_temp = fminf(fmaxf(expf(range_temp), 1e-4f), 1e4f);
  const float inv2s2 = 1.0f / (2.0f * sigma_spatial * sigma_spatial);
  const int rc = cseg_jbu_range_kernel_mma(proj_img, guid_img, n_crops, gh, gw, radius, pos_temp, inv2s2, kern_border, kwidth,
                                           ldk, (cudaStream_t)stream, &sg, CSEG_JBU_FB_RANGE);
  if (rc == 1) CSEG_FAIL(CSEG_EUNSUPPORTED, "jbu_range_kernel_border: radius=%d kwidth=%d not covered (5/128, 3/64)", radius, kwidth);
  return rc;
}

int cseg_jbu_composite_image(const void* kern_img, int ldk, int ih, int iw, int gh, int gw, int radius, void* kc_img,
                             void* tabs_scratch, void* stream) {
  CSEG_REQUIRE(kern_img && kc_img && tabs_scratch && ih > 0 && iw > 0, "jbu_composite_image: null operand / empty image");
  CSEG_REQUIRE(gh >= 2 * CSEG_JBU_FB_COMP + 2 && gw >= 34, "jbu_composite_image: crop region %dx%d too small", gh, gw);
  const int rc = cseg_jbu_composite_image_tc((const bf16*)kern_img, ldk, ih, iw, gh, gw, radius, (bf16*)kc_img, tabs_scratch,
                                             (cudaStream_t)stream);
  if (rc == 1) CSEG_FAIL(CSEG_EUNSUPPORTED, "jbu_composite_image: radius=%d ldk=%d not covered", radius, ldk);
  return rc;
}

int cseg_jbu_apply_shared(const void* src, int n_crops, int h, int w, int C, const void* kern_border, const void* kern_img,
                          const void* kc_img, const cseg_jbu_share* share, int ldk, int radius, void* dst, void* scratch,
                          void* stream) {
  ShareGeom sg;
  if (int rc = share_geom(share, sg, "jbu_apply_shared")) return rc;
  CSEG_REQUIRE(src && kern_border && kern_img && kc_img && dst && scratch && n_crops > 0, "jbu_apply_shared: null operand");
  const int rc = cseg_jbu_apply_shared_tc((const bf16*)src, n_crops, h, w, C, (const bf16*)kern_border, (const bf16*)kern_img,
                                          (const bf16*)kc_img, sg, ldk, radius, (bf16*)dst, scratch, (cudaStream_t)stream);
  if (rc == 1) CSEG_FAIL(CSEG_EUNSUPPORTED, "jbu_apply_shared: shape not covered (C %% 128, radius 3/5, 2h > 2*%d+2, 2w >= 34)", CSEG_JBU_FB_COMP);
  return rc;
}

int cseg_jbu_apply(int dtype, const void* src, int n_crops, int h, int w, int C, const void* kern, int ldk, int radius,
                   void* dst, void* hr_scratch, void* stream) {
  CSEG_REQUIRE(n_crops > 0 && h > 0 && w > 0 && C % 8 == 0, "jbu_apply: C=%d must be a multiple of 8", C);
  CSEG_REQUIRE(radius == 3 || radius == 5, "jbu_apply: radius=%d (3 or 5)", radius);
  CSEG_REQUIRE(2 * h > radius && 2 * w > radius, "jbu_apply: source too small for reflect padding");
  CSEG_REQUIRE(hr_scratch != nullptr, "jbu_apply: hr_scratch (n*2h*2w*C elements) required");
  if (dtype == CSEG_BF16) return launch_apply<bf16>(src, n_crops, h, w, C, kern, ldk, radius, dst, hr_scratch, (cudaStream_t)stream);
  return launch_apply<float>(src, n_crops, h, w, C, kern, ldk, radius, dst, hr_scratch, (cudaStream_t)stream);
}

}  // extern "C"
