// ViT-side kernels that are not GEMMs: input preprocessing, crop -> patch matrix, token assembly,
// LayerNorm, attention (standard and the modified final-block variants), similarity map, outlier
// suppression, CLS normalise + global debias.  Reference: open_clip/transformer.py,
// similarity_enhancement.py, outlier_suppression.py, segmentor.py (cited per kernel).
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// segmentor.py:64-67 (mmseg SegDataPreProcessor): uint8 HWC BGR -> fp32 CHW RGB, (x-mean)/std
// ---------------------------------------------------------------------------------------------
__global__ void preprocess_u8_kernel(const uint8_t* __restrict__ img, int HW, float m0, float m1, float m2, float s0,
                                     float s1, float s2, float* __restrict__ out) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= HW) return;
  const float b = img[3 * i + 0], g = img[3 * i + 1], r = img[3 * i + 2];
  out[i] = (r - m0) / s0;
  out[HW + i] = (g - m1) / s1;
  out[2 * HW + i] = (b - m2) / s2;
}

// ---------------------------------------------------------------------------------------------
// open_clip/transformer.py:560-562 as an im2col gather (the conv has stride == kernel, no bias)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void patchify_kernel(const ImgView img, const int32_t* __restrict__ wins,
                                int n_crops, int gh, int gw, int pad_top, int pad_left, int ps, T* __restrict__ out,
                                int ldo) {
  pdl_grid_sync();
  const long long total = (long long)n_crops * gh * gw * ldo;
  const int kk = 3 * ps * ps;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(idx % ldo);
    const long long row = idx / ldo;
    float v = 0.f;
    if (col < kk) {
      const int p = (int)(row % (gh * gw));
      const int crop = (int)(row / (gh * gw));
      const int c = col / (ps * ps), ky = (col / ps) % ps, kx = col % ps;
      const int cy = (p / gw) * ps + ky - pad_top, cx = (p % gw) * ps + kx - pad_left;
      const int y1 = wins[crop * 4 + 0], x1 = wins[crop * 4 + 1], wh = wins[crop * 4 + 2], ww = wins[crop * 4 + 3];
      if (cy >= 0 && cy < wh && cx >= 0 && cx < ww) v = img_at(img, c, img_row_off(img, y1 + cy), x1 + cx);
    }
    out[idx] = from_f32<T>(v);
  }
}

// bf16 form for ps % 8 == 0: a thread produces 8 consecutive columns (= 8 consecutive pixels of one patch row of one
// channel) and stores them as one 16-byte vector; uint8 inputs are normalised through a 3 x 256 table built per CTA with
// the same IEEE division as img_at (bit-identical values, no division per element).
__global__ void __launch_bounds__(256) patchify_vec8_kernel(const ImgView img, const int32_t* __restrict__ wins, int n_crops,
                                                            int gh, int gw, int pad_top, int pad_left, int ps,
                                                            bf16* __restrict__ out, int ldo) {
  __shared__ float lut[3][256];
  if (img.dtype == CSEG_U8)
    for (int i = threadIdx.x; i < 768; i += blockDim.x) lut[i >> 8][i & 255] = __fdiv_rn((float)(i & 255) - img_mean(img, i >> 8), img_std(img, i >> 8));
  __syncthreads();
  pdl_grid_sync();
  const int g_per_row = ldo >> 3, kk = 3 * ps * ps, P = gh * gw;
  const long long total = (long long)n_crops * P * g_per_row;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(idx % g_per_row) * 8;
    const long long row = idx / g_per_row;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (col < kk) {
      const int p = (int)(row % P), crop = (int)(row / P);
      const int c = col / (ps * ps), ky = (col / ps) % ps, kx = col % ps;
      const int cy = (p / gw) * ps + ky - pad_top, cx = (p % gw) * ps + kx - pad_left;
      const int4 w = *reinterpret_cast<const int4*>(wins + crop * 4);
      if (cy >= 0 && cy < w.z) {
        const long long base = img_row_off(img, w.x + cy) + (long long)img_chan(img, c) * img.stride_c;
        if (img.dtype == CSEG_U8) {
          const uint8_t* src = reinterpret_cast<const uint8_t*>(img.data) + base;
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (cx + e >= 0 && cx + e < w.w) v[e] = lut[c][src[(long long)(w.y + cx + e) * img.stride_x]];
        } else {
          const float* src = reinterpret_cast<const float*>(img.data) + base;
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (cx + e >= 0 && cx + e < w.w) v[e] = src[(long long)(w.y + cx + e) * img.stride_x];
        }
      }
    }
    uint4 pk;
    __nv_bfloat162 t0 = __floats2bfloat162_rn(v[0], v[1]), t1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 t2 = __floats2bfloat162_rn(v[4], v[5]), t3 = __floats2bfloat162_rn(v[6], v[7]);
    pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
    pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
    reinterpret_cast<uint4*>(out)[idx] = pk;
  }
}

// text tower stem (open_clip/model.py:291-293): out[r] = table[idx[r]] + pos[r % L]; pos == nullptr: plain row gather
// (the EOT-token pick of model.py:302-304)
__global__ void gather_rows_kernel(const float* __restrict__ table, const long long* __restrict__ idx,
                                   const float* __restrict__ pos, long long n_rows, int L, int width,
                                   float* __restrict__ out) {
  pdl_grid_sync();
  const long long total = n_rows * width;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / width;
    const int c = (int)(e - r * width);
    float v = table[(size_t)idx[r] * width + c];
    if (pos) v += pos[(size_t)(r % L) * width + c];
    out[e] = v;
  }
}

// open_clip/transformer.py:565-571
__global__ void embed_tokens_kernel(const float* __restrict__ pe, const float* __restrict__ cls,
                                    const float* __restrict__ pos, int n_crops, int L, int width,
                                    float* __restrict__ x) {
  pdl_grid_sync();
  const long long total = (long long)n_crops * L * width;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % width);
    const long long row = idx / width;
    const int t = (int)(row % L);
    const int crop = (int)(row / L);
    const float v = (t == 0) ? cls[c] : pe[((size_t)crop * (L - 1) + (t - 1)) * width + c];
    x[idx] = v + pos[(size_t)t * width + c];
  }
}
// width % 4 == 0: float4 per thread
__global__ void embed_tokens_vec4_kernel(const float4* __restrict__ pe, const float4* __restrict__ cls,
                                         const float4* __restrict__ pos, int n_crops, int L, int w4,
                                         float4* __restrict__ x) {
  pdl_grid_sync();
  const long long total = (long long)n_crops * L * w4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % w4);
    const long long row = idx / w4;
    const int t = (int)(row % L);
    const int crop = (int)(row / L);
    const float4 v = (t == 0) ? __ldg(cls + c) : pe[((size_t)crop * (L - 1) + (t - 1)) * w4 + c];
    const float4 q = __ldg(pos + (size_t)t * w4 + c);
    x[idx] = make_float4(v.x + q.x, v.y + q.y, v.z + q.z, v.w + q.w);
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNormFp32, open_clip/transformer.py:17-23.  One warp per row, two-pass statistics.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void layernorm_kernel(const float* x, int rows, int width, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float eps, T* out) {
  pdl_grid_sync();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* xr = x + (size_t)warp * width;
  float s = 0.f;
  for (int c = lane; c < width; c += 32) s += xr[c];
  const float mean = warp_sum(s) / width;
  float v = 0.f;
  for (int c = lane; c < width; c += 32) {
    const float d = xr[c] - mean;
    v += d * d;
  }
  const float rstd = rsqrtf(warp_sum(v) / width + eps);
  T* orow = out + (size_t)warp * width;
  for (int c = lane; c < width; c += 32) orow[c] = from_f32<T>((xr[c] - mean) * rstd * gamma[c] + beta[c]);
}

// single pass: the row lives in registers (float4 x LNV per lane), for width % 4 == 0 and width <= 128*LNV.
// Grid-stride over rows with the NEXT row of the warp requested before the current one is reduced (the one-row-per-warp form
// ran at 0.64 of the copy bandwidth: every warp paid a full DRAM round trip and a CTA launch for 3 KB of traffic).
constexpr int LNV = 10;
template <typename T, int NV>
__global__ void __launch_bounds__(256) layernorm_reg_kernel(const float* x, int rows, int width,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps, T* out) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  if (warp0 >= rows) return;
  const int nv = width >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  float4 buf[NV], nxt[NV];
  {
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)warp0 * width);
#pragma unroll
    for (int k = 0; k < NV; ++k)
      if (lane + 32 * k < nv) nxt[k] = xr[lane + 32 * k];
  }
  for (int row = warp0; row < rows; row += nwarps) {
#pragma unroll
    for (int k = 0; k < NV; ++k) buf[k] = nxt[k];
    if (row + nwarps < rows) {
      const float4* xn = reinterpret_cast<const float4*>(x + (size_t)(row + nwarps) * width);
#pragma unroll
      for (int k = 0; k < NV; ++k)
        if (lane + 32 * k < nv) nxt[k] = xn[lane + 32 * k];
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k)
      if (lane + 32 * k < nv) s += (buf[k].x + buf[k].y) + (buf[k].z + buf[k].w);
    const float mean = warp_sum(s) / width;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      if (lane + 32 * k < nv) {
        const float a = buf[k].x - mean, b = buf[k].y - mean, c = buf[k].z - mean, d = buf[k].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / width + eps);
    T* orow = out + (size_t)row * width;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int v = lane + 32 * k;
      if (v < nv) {
        const float4 g = __ldg(g4 + v), b = __ldg(b4 + v);
        const float o0 = (buf[k].x - mean) * rstd * g.x + b.x, o1 = (buf[k].y - mean) * rstd * g.y + b.y;
        const float o2 = (buf[k].z - mean) * rstd * g.z + b.z, o3 = (buf[k].w - mean) * rstd * g.w + b.w;
        if (sizeof(T) == 4) {
          reinterpret_cast<float4*>(orow)[v] = make_float4(o0, o1, o2, o3);
        } else {
          __nv_bfloat162 lo = __floats2bfloat162_rn(o0, o1), hi = __floats2bfloat162_rn(o2, o3);
          uint2 u;
          u.x = *reinterpret_cast<uint32_t*>(&lo);
          u.y = *reinterpret_cast<uint32_t*>(&hi);
          reinterpret_cast<uint2*>(orow)[v] = u;
        }
      }
    }
  }
}

// token assembly (transformer.py:565-571) + ln_pre (:574) in one pass: x[crop, t] = LN((t == 0 ? cls : pe[crop, t-1]) + pos[t]).
// One warp per token row, the row in registers; same arithmetic order as embed_tokens followed by layernorm_reg.
__global__ void __launch_bounds__(256) embed_ln_kernel(const float4* __restrict__ pe, const float4* __restrict__ cls,
                                                       const float4* __restrict__ pos, int n_crops, int L, int width,
                                                       const float4* __restrict__ gamma, const float4* __restrict__ beta,
                                                       float eps, float4* __restrict__ x) {
  pdl_grid_sync();
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= (long long)n_crops * L) return;
  const int nv = width >> 2;
  const int t = (int)(row % L);
  const long long crop = row / L;
  const float4* src = (t == 0) ? cls : pe + ((size_t)crop * (L - 1) + (t - 1)) * nv;
  const float4* pr = pos + (size_t)t * nv;
  float4 buf[LNV];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < LNV; ++k) {
    const int v = lane + 32 * k;
    if (v < nv) {
      const float4 a = src[v], q = __ldg(pr + v);
      buf[k] = make_float4(a.x + q.x, a.y + q.y, a.z + q.z, a.w + q.w);
      s += (buf[k].x + buf[k].y) + (buf[k].z + buf[k].w);
    }
  }
  const float mean = warp_sum(s) / width;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < LNV; ++k) {
    const int v = lane + 32 * k;
    if (v < nv) {
      const float a = buf[k].x - mean, b = buf[k].y - mean, c = buf[k].z - mean, d = buf[k].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / width + eps);
  float4* orow = x + (size_t)row * nv;
#pragma unroll
  for (int k = 0; k < LNV; ++k) {
    const int v = lane + 32 * k;
    if (v < nv) {
      const float4 g = __ldg(gamma + v), b = __ldg(beta + v);
      orow[v] = make_float4((buf[k].x - mean) * rstd * g.x + b.x, (buf[k].y - mean) * rstd * g.y + b.y,
                            (buf[k].z - mean) * rstd * g.z + b.z, (buf[k].w - mean) * rstd * g.w + b.w);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Attention.  One CTA per (crop, head); Q, K, V of the head live in shared memory (type T); each warp
// owns query rows i = warp, warp+8, ...; a lane owns keys j = lane + 32 t.  All softmax arithmetic
// is fp32.  Reference: nn.MultiheadAttention (open_clip/transformer.py:204,218-232) for mode STD and
// VisionTransformer.custom_attn (:822-940) for the others.
// ---------------------------------------------------------------------------------------------
constexpr int ATT_WARPS = 8;
constexpr int ATT_JMAX = 10;  // L <= 320

template <typename T> struct Pair;
template <> struct Pair<float> {
  static __device__ __forceinline__ float2 ld(const float* p) { return make_float2(p[0], p[1]); }
  static __device__ __forceinline__ void st(float* p, float a, float b) { p[0] = a; p[1] = b; }
};
template <> struct Pair<bf16> {
  static __device__ __forceinline__ float2 ld(const bf16* p) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
  }
  static __device__ __forceinline__ void st(bf16* p, float a, float b) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
  }
};

template <typename T, int HD>
__device__ __forceinline__ float dot_row(const float (&r)[HD], const T* row) {
  float acc = 0.f;
#pragma unroll
  for (int d = 0; d < HD; d += 2) {
    const float2 v = Pair<T>::ld(row + d);
    acc = fmaf(r[d], v.x, acc);
    acc = fmaf(r[d + 1], v.y, acc);
  }
  return acc;
}

template <typename T, int HD>
__device__ __forceinline__ void load_row(float (&r)[HD], const T* row) {
#pragma unroll
  for (int d = 0; d < HD; d += 2) {
    const float2 v = Pair<T>::ld(row + d);
    r[d] = v.x;
    r[d + 1] = v.y;
  }
}

// softmax over the keys a warp holds (s[t] for key lane+32t, invalid keys hold -inf)
__device__ __forceinline__ void warp_softmax(float (&s)[ATT_JMAX], int nj) {
  float m = -INFINITY;
#pragma unroll
  for (int t = 0; t < ATT_JMAX; ++t)
    if (t < nj) m = fmaxf(m, s[t]);
  m = warp_max(m);
  float sum = 0.f;
#pragma unroll
  for (int t = 0; t < ATT_JMAX; ++t)
    if (t < nj) {
      s[t] = __expf(s[t] - m);
      sum += s[t];
    }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
#pragma unroll
  for (int t = 0; t < ATT_JMAX; ++t)
    if (t < nj) s[t] *= inv;
}

template <typename T, int HD>
__global__ void __launch_bounds__(ATT_WARPS * 32) attention_kernel(const T* __restrict__ qkv, int L, int heads,
                                                                   int mode, const float* __restrict__ simmap,
                                                                   float simw, T* __restrict__ out,
                                                                   float* __restrict__ stats) {
  pdl_grid_sync();
  constexpr int LDS = HD + (sizeof(T) == 2 ? 2 : 1);
  extern __shared__ __align__(16) uint8_t att_smem[];
  T* Qs = reinterpret_cast<T*>(att_smem);
  T* Ks = Qs + (size_t)L * LDS;
  T* Vs = Ks + (size_t)L * LDS;
  float* Pb = reinterpret_cast<float*>(att_smem + (((size_t)3 * L * LDS * sizeof(T) + 15) & ~(size_t)15));
  const int crop = blockIdx.x / heads, head = blockIdx.x % heads;
  const int width = heads * HD;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float scale = rsqrtf((float)HD);
  const int P = L - 1;

  for (int idx = tid; idx < L * (HD / 2); idx += blockDim.x) {
    const int row = idx / (HD / 2), c = (idx % (HD / 2)) * 2;
    const T* src = qkv + ((size_t)crop * L + row) * 3 * width + head * HD + c;
    const float2 q = Pair<T>::ld(src), k = Pair<T>::ld(src + width), v = Pair<T>::ld(src + 2 * width);
    Pair<T>::st(Qs + row * LDS + c, q.x, q.y);
    Pair<T>::st(Ks + row * LDS + c, k.x, k.y);
    Pair<T>::st(Vs + row * LDS + c, v.x, v.y);
  }
  __syncthreads();

  const int nj = (L + 31) / 32;
  float* pb = Pb + warp * (ATT_JMAX * 32);
  for (int i = warp; i < L; i += ATT_WARPS) {
    float p[ATT_JMAX];
    if (mode == CSEG_ATTN_MASKCLIP) {
#pragma unroll
      for (int t = 0; t < ATT_JMAX; ++t) p[t] = (lane + 32 * t == i) ? 1.f : 0.f;
    } else {
      float qi[HD], ki[HD];
      load_row<T, HD>(qi, Qs + i * LDS);
      const bool need_k = (mode == CSEG_ATTN_EXPERIMENTAL || mode == CSEG_ATTN_SCLIP || mode == CSEG_ATTN_SFP ||
                           mode == CSEG_ATTN_SEGEARTH);
      if (need_k) load_row<T, HD>(ki, Ks + i * LDS);
      // similarity-map row (zero CLS row / column), similarity_enhancement.py:104-122
      float madd[ATT_JMAX];
#pragma unroll
      for (int t = 0; t < ATT_JMAX; ++t) {
        const int j = lane + 32 * t;
        madd[t] = (simmap != nullptr && i >= 1 && j >= 1 && j < L)
                      ? simw * simmap[((size_t)crop * P + (i - 1)) * P + (j - 1)]
                      : 0.f;
      }
      float s1[ATT_JMAX], s2[ATT_JMAX];
#pragma unroll
      for (int t = 0; t < ATT_JMAX; ++t) {
        const int j = lane + 32 * t;
        s1[t] = -INFINITY;
        s2[t] = -INFINITY;
        if (t < nj && j < L) {
          if (mode == CSEG_ATTN_CAUSAL) {
            if (j <= i) s1[t] = dot_row<T, HD>(qi, Ks + j * LDS) * scale;
          } else if (mode == CSEG_ATTN_STD || mode == CSEG_ATTN_VANILLA) {
            s1[t] = dot_row<T, HD>(qi, Ks + j * LDS) * scale;
          } else {
            s1[t] = dot_row<T, HD>(qi, Qs + j * LDS) * scale;
            if (need_k) s2[t] = dot_row<T, HD>(ki, Ks + j * LDS) * scale;
          }
        }
      }
      if (mode == CSEG_ATTN_STD || mode == CSEG_ATTN_CAUSAL) {
        warp_softmax(s1, nj);
      } else if (mode == CSEG_ATTN_VANILLA || mode == CSEG_ATTN_CLEARCLIP) {
#pragma unroll
        for (int t = 0; t < ATT_JMAX; ++t) s1[t] += madd[t];
        warp_softmax(s1, nj);
      } else if (mode == CSEG_ATTN_SFP) {
#pragma unroll
        for (int t = 0; t < ATT_JMAX; ++t) s1[t] = 0.5f * (s1[t] + s2[t]) + madd[t];
        warp_softmax(s1, nj);
      } else if (mode == CSEG_ATTN_EXPERIMENTAL) {
#pragma unroll
        for (int t = 0; t < ATT_JMAX; ++t) s1[t] = s2[t] + s1[t];  // kk + qq, :899
        warp_softmax(s1, nj);
#pragma unroll
        for (int t = 0; t < ATT_JMAX; ++t)  // enhance_attention on the probabilities, :901
          s1[t] = (lane + 32 * t < L) ? s1[t] + madd[t] : -INFINITY;
        warp_softmax(s1, nj);                       // second softmax is unconditional, :902
      } else {                                      // SCLIP / SEGEARTH
#pragma unroll
        for (int t = 0; t < ATT_JMAX; ++t) {
          s1[t] += madd[t];
          s2[t] += madd[t];
        }
        warp_softmax(s1, nj);
        warp_softmax(s2, nj);
#pragma unroll
        for (int t = 0; t < ATT_JMAX; ++t) s1[t] += s2[t];
        if (mode == CSEG_ATTN_SEGEARTH) {
          load_row<T, HD>(ki, Vs + i * LDS);
#pragma unroll
          for (int t = 0; t < ATT_JMAX; ++t) {
            const int j = lane + 32 * t;
            s2[t] = (t < nj && j < L) ? dot_row<T, HD>(ki, Vs + j * LDS) * scale + madd[t] : -INFINITY;
          }
          warp_softmax(s2, nj);
#pragma unroll
          for (int t = 0; t < ATT_JMAX; ++t) s1[t] += s2[t];
        }
      }
#pragma unroll
      for (int t = 0; t < ATT_JMAX; ++t)
        p[t] = (lane + 32 * t < L && !(mode == CSEG_ATTN_CAUSAL && lane + 32 * t > i)) ? s1[t] : 0.f;
    }
    if (stats != nullptr) {  // outlier_suppression.py:46-49: only P[0,1+i] and P[1+i,1+i] are consumed
      float* st = stats + ((size_t)(crop * heads + head) * 2) * P;
#pragma unroll
      for (int t = 0; t < ATT_JMAX; ++t) {
        const int j = lane + 32 * t;
        if (j >= 1 && j < L) {
          if (i == 0) st[j - 1] = p[t];
          if (j == i) st[P + i - 1] = p[t];
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int t = 0; t < ATT_JMAX; ++t)
      if (t < nj) pb[lane + 32 * t] = p[t];
    __syncwarp();
    for (int dp = lane; dp < HD / 2; dp += 32) {
      float a0 = 0.f, a1 = 0.f;
      for (int j = 0; j < L; ++j) {
        const float pj = pb[j];
        const float2 v = Pair<T>::ld(Vs + j * LDS + 2 * dp);
        a0 = fmaf(pj, v.x, a0);
        a1 = fmaf(pj, v.y, a1);
      }
      Pair<T>::st(out + ((size_t)crop * L + i) * width + head * HD + 2 * dp, a0, a1);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Attention for long token sequences (whole-image inference, segmentor.py:470-471: L = H W / ps^2 + 1 is unbounded;
// slide_crop 336 with ViT-L/14: L = 577).  Same formulas as attention_kernel, any L: a warp owns one query row at a
// time and keeps that row's scores in shared memory (L floats per warp) instead of registers; K / V / Q rows of the
// keys are read from global memory (the warps of a CTA walk the same rows: L1 / L2 hits).  Every mode is a sum of one
// to three softmax terms, each over a sum of one or two score products; a term's unnormalised probabilities are
// contracted with V as soon as they exist, so one score row per warp is enough.
// ---------------------------------------------------------------------------------------------
template <typename T, int HD>
__global__ void __launch_bounds__(256) attention_long_kernel(const T* __restrict__ qkv, int L, int heads, int mode,
                                                             const float* __restrict__ simmap, float simw, T* __restrict__ out,
                                                             float* __restrict__ stats, int rows_per_cta) {
  pdl_grid_sync();
  extern __shared__ __align__(16) uint8_t att_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float* sc = reinterpret_cast<float*>(att_smem) + (size_t)warp * L;
  const int nrb = (L + rows_per_cta - 1) / rows_per_cta;
  const int item = blockIdx.x / nrb, rb = blockIdx.x - item * nrb;
  const int crop = item / heads, head = item - crop * heads, width = heads * HD, P = L - 1;
  const float scale = rsqrtf((float)HD);
  const T* base = qkv + (size_t)crop * L * 3 * width + head * HD;      // + row * 3 width (+ width: K, + 2 width: V)
  const size_t rs = (size_t)3 * width;
  constexpr int DPL = (HD + 31) / 32;                                   // output dimensions per lane: d = lane + 32 u
  const int i_end = min(L, (rb + 1) * rows_per_cta);
  for (int i = rb * rows_per_cta + warp; i < i_end; i += nwarps) {
    float o[DPL];
#pragma unroll
    for (int u = 0; u < DPL; ++u) o[u] = 0.f;
    if (mode == CSEG_ATTN_MASKCLIP) {                                   // identity attention: out_i = v_i
#pragma unroll
      for (int u = 0; u < DPL; ++u)
        if (lane + 32 * u < HD) o[u] = to_f32(base[(size_t)i * rs + 2 * width + lane + 32 * u]);
    } else {
      const int nterms = mode == CSEG_ATTN_SCLIP ? 2 : (mode == CSEG_ATTN_SEGEARTH ? 3 : 1);
      const bool use_m = simmap != nullptr && mode != CSEG_ATTN_STD && mode != CSEG_ATTN_CAUSAL && i >= 1;
      const float* mrow = use_m ? simmap + ((size_t)crop * P + (i - 1)) * P : nullptr;   // M_pad[i][j] = M[i-1][j-1], j >= 1
      for (int term = 0; term < nterms; ++term) {
        // ---- scores of this term into sc[] ----
        const int nsub = (mode == CSEG_ATTN_SFP || mode == CSEG_ATTN_EXPERIMENTAL) ? 2 : 1;
        float m = -INFINITY;
        for (int sub = 0; sub < nsub; ++sub) {
          // x_i . Y_j: STD / vanilla / causal q.K; ClearCLIP q.Q; SFP, Experimental q.Q then k.K; SCLIP / SegEarth terms q.Q, k.K, v.V
          int xo, yo;
          if (mode == CSEG_ATTN_STD || mode == CSEG_ATTN_VANILLA || mode == CSEG_ATTN_CAUSAL) { xo = 0; yo = width; }
          else if (mode == CSEG_ATTN_SCLIP || mode == CSEG_ATTN_SEGEARTH) { xo = yo = term * width; }
          else { xo = yo = sub * width; }
          const float w = mode == CSEG_ATTN_SFP ? 0.5f * scale : scale;
          float xi[HD];
          load_row<T, HD>(xi, base + (size_t)i * rs + xo);
          const bool last = sub == nsub - 1;
          const bool add_m = last && mode != CSEG_ATTN_EXPERIMENTAL;       // Experimental adds M to the probabilities
          for (int j = lane; j < L; j += 32) {
            float v = dot_row<T, HD>(xi, base + (size_t)j * rs + yo) * w;
            if (sub > 0) v += sc[j];
            if (add_m && use_m && j >= 1) v += simw * mrow[j - 1];
            if (mode == CSEG_ATTN_CAUSAL && j > i) v = -INFINITY;
            sc[j] = v;
            if (last) m = fmaxf(m, v);
          }
        }
        m = warp_max(m);
        float sum = 0.f;
        for (int j = lane; j < L; j += 32) {
          const float e = __expf(sc[j] - m);
          sc[j] = e;
          sum += e;
        }
        sum = warp_sum(sum);
        float inv = 1.0f / sum;
        if (mode == CSEG_ATTN_EXPERIMENTAL) {                               // second softmax over p1 + w M (transformer.py:901-902)
          float m2 = -INFINITY;
          for (int j = lane; j < L; j += 32) {
            float v = sc[j] * inv;
            if (use_m && j >= 1) v += simw * mrow[j - 1];
            sc[j] = v;
            m2 = fmaxf(m2, v);
          }
          m2 = warp_max(m2);
          sum = 0.f;
          for (int j = lane; j < L; j += 32) {
            const float e = __expf(sc[j] - m2);
            sc[j] = e;
            sum += e;
          }
          sum = warp_sum(sum);
          inv = 1.0f / sum;
        }
        if (stats != nullptr) {   // outlier_suppression.py:46-49: only P[0,1+i] and P[1+i,1+i] are consumed (STD only)
          float* st = stats + ((size_t)(crop * heads + head) * 2) * P;
          if (i == 0)
            for (int j = 1 + lane; j < L; j += 32) st[j - 1] = sc[j] * inv;
          else if (lane == (i & 31)) st[P + i - 1] = sc[i] * inv;
        }
        __syncwarp();
        // ---- o += (1 / sum) * sum_j e_j v_j ----
        float a[DPL];
#pragma unroll
        for (int u = 0; u < DPL; ++u) a[u] = 0.f;
        const T* vb = base + 2 * width;
        const int jn = mode == CSEG_ATTN_CAUSAL ? i + 1 : L;
#pragma unroll 4
        for (int j = 0; j < jn; ++j) {
          const float pj = sc[j];
#pragma unroll
          for (int u = 0; u < DPL; ++u)
            if (lane + 32 * u < HD) a[u] = fmaf(pj, to_f32(vb[(size_t)j * rs + lane + 32 * u]), a[u]);
        }
#pragma unroll
        for (int u = 0; u < DPL; ++u) o[u] = fmaf(a[u], inv, o[u]);
        __syncwarp();
      }
    }
#pragma unroll
    for (int u = 0; u < DPL; ++u)
      if (lane + 32 * u < HD) out[((size_t)crop * L + i) * width + head * HD + lane + 32 * u] = from_f32<T>(o[u]);
  }
}

// ---------------------------------------------------------------------------------------------
// similarity_enhancement.py:37-66.  inv_norm then a 32x32-tiled fp32 dot-product kernel.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) simmap_kernel(const float* __restrict__ x, int L, int width, float inv_temp,
                                                     int keep_diag, float* __restrict__ sim) {
  pdl_grid_sync();
  // 64 x 64 tile per CTA, 4 x 4 outputs per thread, operands staged k-major so that a thread reads its four rows /
  // columns as one float4 (broadcast across the 16 threads that share them)
  constexpr int TS = 64, KC = 16, LDS_ = TS + 4;
  __shared__ __align__(16) float As[KC][LDS_], Bs[KC][LDS_];
  if (blockIdx.x < blockIdx.y) return;                   // M is symmetric: tiles on / above the diagonal write both halves
  const int P = L - 1, crop = blockIdx.z;
  const int i0 = blockIdx.y * TS, j0 = blockIdx.x * TS;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float* xb = x + ((size_t)crop * L + 1) * width;  // skip CLS
  float acc[4][4] = {};
  float na[4] = {}, nb[4] = {};   // squared norms of the rows this thread touches (F.normalize)
  for (int k0 = 0; k0 < width; k0 += KC) {
#pragma unroll
    for (int n = 0; n < TS * KC / 256; ++n) {
      const int e = threadIdx.x + n * 256, r = e / KC, k = e % KC;
      As[k][r] = (i0 + r < P && k0 + k < width) ? xb[(size_t)(i0 + r) * width + k0 + k] : 0.f;
      Bs[k][r] = (j0 + r < P && k0 + k < width) ? xb[(size_t)(j0 + r) * width + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        na[u] = fmaf(a[u], a[u], na[u]);
        nb[u] = fmaf(b[u], b[u], nb[u]);
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(a[u], b[v], acc[u][v]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int i = i0 + ty * 4 + a, j = j0 + tx * 4 + b;
      if (i < P && j < P) {
        const float ia = 1.0f / fmaxf(sqrtf(na[a]), 1e-12f), ib = 1.0f / fmaxf(sqrtf(nb[b]), 1e-12f);
        float v = acc[a][b] * (ia * ib) * inv_temp;       // (ia * ib): the same value for (i, j) and (j, i)
        if (!keep_diag && i == j) v = 0.f;
        sim[((size_t)crop * P + i) * P + j] = v;
        if (blockIdx.x != blockIdx.y) sim[((size_t)crop * P + j) * P + i] = v;
      }
    }
}

// Tensor-core form of the similarity map: rows are L2-normalised (F.normalize, eps 1e-12) and written as a [hi | lo] bf16
// split (hi = bf16(v), lo = bf16(v - hi)): hi.hi + hi.lo + lo.hi reproduces the fp32 dot product to ~2^-17 relative.
// One warp per row, the row in registers.
__global__ void __launch_bounds__(256) simmap_split_kernel(const float* __restrict__ x, long long rows, int width,
                                                           bf16* __restrict__ xs) {
  pdl_grid_sync();
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int nv = width >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * width);
  float4 buf[LNV];
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < LNV; ++k) {
    const int v = lane + 32 * k;
    if (v < nv) {
      buf[k] = xr[v];
      q += (buf[k].x * buf[k].x + buf[k].y * buf[k].y) + (buf[k].z * buf[k].z + buf[k].w * buf[k].w);
    }
  }
  const float inv = 1.0f / fmaxf(sqrtf(warp_sum(q)), 1e-12f);
  uint2* hi = reinterpret_cast<uint2*>(xs + (size_t)row * 2 * width);
  uint2* lo = reinterpret_cast<uint2*>(xs + (size_t)row * 2 * width + width);
#pragma unroll
  for (int k = 0; k < LNV; ++k) {
    const int v = lane + 32 * k;
    if (v < nv) {
      const float a[4] = {buf[k].x * inv, buf[k].y * inv, buf[k].z * inv, buf[k].w * inv};
      bf16 h[4], l[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        h[e] = __float2bfloat16_rn(a[e]);
        l[e] = __float2bfloat16_rn(a[e] - __bfloat162float(h[e]));
      }
      hi[v] = *reinterpret_cast<const uint2*>(h);
      lo[v] = *reinterpret_cast<const uint2*>(l);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// outlier_suppression.py:15-61,115-214, zero host syncs, two kernels:
//   plan  (one CTA per crop): ratio_i = mean_h P[0,1+i] / (mean_h P[1+i,1+i] + 1e-8); top-k by repeated
//          argmax (ties -> lowest index); per (outlier, neighbour) cosine on the ORIGINAL map; softmax
//          weights; and the "owner" of every cell under the reference's sequential write order
//          (neighbours in (i, j) order, last writer wins, cells equal to the outlier skipped; outliers last).
//   apply (one CTA per cell): writes the output row from the ORIGINAL map and the plan -- out of place, so no
//          ordering between CTAs is needed:  none -> copy;  neighbour (i,j) -> nb - clamp(sim*temp,0,1) * o_i;
//          outlier i -> sum_j w_ij * nb_ij.
// plan layout per crop (32-bit words): idx[k] | nb[8k] | sim[8k] | w[8k] | owner[P]
// ---------------------------------------------------------------------------------------------
constexpr int OS_THREADS = 256;
constexpr int OS_MAXK = 64;

__global__ void __launch_bounds__(OS_THREADS) outlier_plan_kernel(const float* __restrict__ y, int L, int width, int grid,
                                                                  const float* __restrict__ stats, int heads, int top_k,
                                                                  int* __restrict__ plan, int32_t* __restrict__ idx_out) {
  pdl_grid_sync();
  extern __shared__ float os_smem[];
  const int P = L - 1, crop = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* ratio = os_smem;                       // [P]
  __shared__ int s_idx[OS_MAXK];
  __shared__ float s_sim[OS_MAXK][8];
  __shared__ int s_nb[OS_MAXK][8];
  __shared__ float red_v[OS_THREADS / 32];
  __shared__ int red_i[OS_THREADS / 32];
  const int k = min(top_k, P);
  int* pl = plan + (size_t)crop * (25 * top_k + P);
  int* p_idx = pl;
  int* p_nb = pl + top_k;
  float* p_sim = reinterpret_cast<float*>(pl + 9 * top_k);
  float* p_w = reinterpret_cast<float*>(pl + 17 * top_k);
  int* p_owner = pl + 25 * top_k;

  for (int i = tid; i < P; i += OS_THREADS) {
    float c = 0.f, d = 0.f;
    for (int h = 0; h < heads; ++h) {
      const float* st = stats + ((size_t)(crop * heads + h) * 2) * P;
      c += st[i];
      d += st[P + i];
    }
    ratio[i] = (c / heads) / (d / heads + 1e-8f);
    p_owner[i] = -1;
  }
  __syncthreads();
  for (int r = 0; r < k; ++r) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = tid; i < P; i += OS_THREADS) {
      const float v = ratio[i];
      if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { red_v[warp] = bv; red_i[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < OS_THREADS / 32; ++w)
        if (red_v[w] > bv || (red_v[w] == bv && red_i[w] < bi)) { bv = red_v[w]; bi = red_i[w]; }
      s_idx[r] = bi;
      ratio[bi] = -INFINITY;
      p_idx[r] = bi;
      if (idx_out) idx_out[crop * top_k + r] = bi;
    }
    __syncthreads();
  }
  const float* yb = y + ((size_t)crop * L + 1) * width;  // patch rows
  const int offs_y[8] = {-1, -1, -1, 0, 0, 1, 1, 1};
  const int offs_x[8] = {-1, 0, 1, -1, 1, -1, 0, 1};
  for (int pr = warp; pr < k * 8; pr += OS_THREADS / 32) {   // one warp per (outlier, neighbour)
    const int i = pr >> 3, j = pr & 7;
    const int oy = s_idx[i] / grid, ox = s_idx[i] % grid;
    const int ny = min(max(oy + offs_y[j], 0), grid - 1), nx = min(max(ox + offs_x[j], 0), grid - 1);
    const float* o = yb + (size_t)s_idx[i] * width;
    const float* nb = yb + (size_t)(ny * grid + nx) * width;
    float dot = 0.f, no = 0.f, nn = 0.f;
    for (int c = lane; c < width; c += 32) {
      const float a = o[c], b = nb[c];
      dot = fmaf(a, b, dot);
      no = fmaf(a, a, no);
      nn = fmaf(b, b, nn);
    }
    dot = warp_sum(dot); no = warp_sum(no); nn = warp_sum(nn);
    if (lane == 0) {
      s_sim[i][j] = dot / (fmaxf(sqrtf(no), 1e-12f) * fmaxf(sqrtf(nn), 1e-12f));
      s_nb[i][j] = ny * grid + nx;
    }
  }
  __syncthreads();
  for (int i = tid; i < k; i += OS_THREADS) {   // softmax(clamp(1 - sim, 0)) over the 8 neighbours
    float w[8], m = -INFINITY, sm = 0.f;
    for (int j = 0; j < 8; ++j) { w[j] = fmaxf(1.0f - s_sim[i][j], 0.f); m = fmaxf(m, w[j]); }
    for (int j = 0; j < 8; ++j) { w[j] = expf(w[j] - m); sm += w[j]; }
    for (int j = 0; j < 8; ++j) {
      p_w[i * 8 + j] = w[j] / sm;
      p_sim[i * 8 + j] = s_sim[i][j];
      p_nb[i * 8 + j] = s_nb[i][j];
    }
  }
  if (tid == 0) {                                // sequential write order of the reference (:204-212)
    for (int i = 0; i < k; ++i)
      for (int j = 0; j < 8; ++j)
        if (s_nb[i][j] != s_idx[i]) p_owner[s_nb[i][j]] = 1000 + i * 8 + j;
    for (int i = 0; i < k; ++i) p_owner[s_idx[i]] = i;
  }
}

__global__ void __launch_bounds__(128) outlier_apply_kernel(const float* __restrict__ y, float* __restrict__ out, int L,
                                                            int width, int top_k, float ctemp,
                                                            const int* __restrict__ plan) {
  pdl_grid_sync();
  const int P = L - 1, crop = blockIdx.x / L, t = blockIdx.x % L;
  const float* src = y + ((size_t)crop * L + t) * width;
  float* dst = out + ((size_t)crop * L + t) * width;
  if (t == 0) {                                  // CLS token bypasses the module (transformer.py:727,741)
    for (int c = threadIdx.x; c < width; c += 128) dst[c] = src[c];
    return;
  }
  const int* pl = plan + (size_t)crop * (25 * top_k + P);
  const int* p_idx = pl;
  const int* p_nb = pl + top_k;
  const float* p_sim = reinterpret_cast<const float*>(pl + 9 * top_k);
  const float* p_w = reinterpret_cast<const float*>(pl + 17 * top_k);
  const int owner = pl[25 * top_k + (t - 1)];
  const float* yb = y + ((size_t)crop * L + 1) * width;
  if (owner < 0) {
    for (int c = threadIdx.x; c < width; c += 128) dst[c] = src[c];
  } else if (owner >= 1000) {
    const int ij = owner - 1000, i = ij >> 3;
    const float strength = fminf(fmaxf(p_sim[ij] * ctemp, 0.f), 1.f);
    const float* o = yb + (size_t)p_idx[i] * width;
    for (int c = threadIdx.x; c < width; c += 128) dst[c] = src[c] - o[c] * strength;
  } else {
    const int i = owner;
    for (int c = threadIdx.x; c < width; c += 128) {
      float rep = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) rep = fmaf(yb[(size_t)p_nb[i * 8 + j] * width + c], p_w[i * 8 + j], rep);
      dst[c] = rep;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// segmentor.py:309-336: CLS L2-normalise (in place in the reference) + similarity-weighted debias
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void cls_debias_kernel(const float* __restrict__ tok, int n_crops, int L, int D, float factor,
                                  T* __restrict__ feats, int ldf, int rows, float* __restrict__ cls_unit) {
  pdl_grid_sync();
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int P = rows;                       // output rows per crop: L-1 patch tokens, then zero padding
  if (gw >= n_crops * (rows + 1)) return;
  const int crop = gw / (rows + 1), t = gw % (rows + 1);
  if (t >= L) {                             // padding row
    T* o = feats + ((size_t)crop * P + (t - 1)) * ldf;
    for (int c = lane; c < D; c += 32) o[c] = from_f32<T>(0.f);
    return;
  }
  const float* cls = tok + (size_t)crop * L * D;
  float cs = 0.f;
  for (int c = lane; c < D; c += 32) cs = fmaf(cls[c], cls[c], cs);
  const float cinv = 1.0f / sqrtf(warp_sum(cs));
  if (t == 0) {
    if (cls_unit)
      for (int c = lane; c < D; c += 32) cls_unit[(size_t)crop * D + c] = cls[c] * cinv;
    return;
  }
  const float* f = tok + ((size_t)crop * L + t) * D;
  T* o = feats + ((size_t)crop * P + (t - 1)) * ldf;
  if (factor == 0.f) {
    for (int c = lane; c < D; c += 32) o[c] = from_f32<T>(f[c]);
    return;
  }
  // cls_unit is re-normalised in the reference (segmentor.py:325); |cls_unit| = 1 up to rounding
  float fs = 0.f, dot = 0.f, c2 = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float cu = cls[c] * cinv;
    fs = fmaf(f[c], f[c], fs);
    dot = fmaf(f[c], cu, dot);
    c2 = fmaf(cu, cu, c2);
  }
  fs = warp_sum(fs); dot = warp_sum(dot); c2 = warp_sum(c2);
  const float sim = dot / (sqrtf(fs) * sqrtf(c2));
  const float wf = sim * factor;
  for (int c = lane; c < D; c += 32) o[c] = from_f32<T>(f[c] - cls[c] * cinv * wf);
}

}  // namespace

int cseg_attention_mma(const bf16* qkv, int n_crops, int L, int heads, int head_dim, int mode, const float* simmap,
                       float simw, bf16* out, float* stats, cudaStream_t st);

extern "C" {

int cseg_preprocess_u8(const uint8_t* img, int H, int W, const float mean[3], const float std_[3], float* out,
                       void* stream) {
  CSEG_REQUIRE(H > 0 && W > 0, "preprocess: empty image");
  const int HW = H * W;
  cseg_launch(preprocess_u8_kernel, dim3(cdiv(HW, 256)), dim3(256), 0, (cudaStream_t)stream, img, HW, mean[0], mean[1], mean[2], std_[0],
                                                                        std_[1], std_[2], out);
  CSEG_LAUNCH_CHECK("preprocess_u8");
  return 0;
}

int cseg_patchify(const cseg_image* img_desc, const int32_t* windows, int n_crops, int crop_h, int crop_w,
                  int pad_top, int pad_left, int ps, int out_dtype, void* out, int ldo, void* stream) {
  ImgView img;
  CSEG_REQUIRE(make_img_view(img_desc, img) == 0, "patchify: bad image descriptor");
  CSEG_REQUIRE(n_crops > 0 && ps > 0 && crop_h % ps == 0 && crop_w % ps == 0,
               "patchify: crop %dx%d must be a multiple of the patch size %d", crop_h, crop_w, ps);
  CSEG_REQUIRE(ldo >= 3 * ps * ps, "patchify: ldo=%d < %d", ldo, 3 * ps * ps);
  const int gh = crop_h / ps, gw = crop_w / ps;
  const long long total = (long long)n_crops * gh * gw * ldo;
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)sm_count() * 32);
  if (out_dtype == CSEG_BF16 && ps % 8 == 0 && ldo % 8 == 0 && ((uintptr_t)out & 15) == 0) {
    const long long groups = total / 8;
    const int vb = (int)std::min<long long>((groups + 255) / 256, (long long)sm_count() * 16);
    cseg_launch(patchify_vec8_kernel, dim3(vb), dim3(256), 0, (cudaStream_t)stream, img, windows, n_crops, gh, gw, pad_top,
                                                              pad_left, ps, (bf16*)out, ldo);
  } else if (out_dtype == CSEG_BF16)
    cseg_launch(patchify_kernel<bf16>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, img, windows, n_crops, gh, gw, pad_top,
                                                                    pad_left, ps, (bf16*)out, ldo);
  else
    cseg_launch(patchify_kernel<float>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, img, windows, n_crops, gh, gw, pad_top,
                                                                     pad_left, ps, (float*)out, ldo);
  CSEG_LAUNCH_CHECK("patchify");
  return 0;
}

int cseg_gather_rows(const float* table, const long long* idx, const float* pos, long long n_rows, int L, int width,
                     float* out, void* stream) {
  CSEG_REQUIRE(table && idx && out && n_rows > 0 && width > 0 && L > 0, "gather_rows: bad arguments");
  const int blocks = (int)std::min<long long>((n_rows * width + 255) / 256, (long long)sm_count() * 16);
  cseg_launch(gather_rows_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, table, idx, pos, n_rows, L, width, out);
  CSEG_LAUNCH_CHECK("gather_rows");
  return 0;
}

int cseg_embed_tokens(const float* pe, const float* cls, const float* pos, int n_crops, int L, int width, float* x,
                      void* stream) {
  CSEG_REQUIRE(n_crops > 0 && L > 1 && width > 0, "embed_tokens: bad shape");
  const long long total = (long long)n_crops * L * width;
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)sm_count() * 32);
  if (width % 4 == 0 && (((uintptr_t)pe | (uintptr_t)cls | (uintptr_t)pos | (uintptr_t)x) & 15) == 0) {
    const int vb = (int)std::min<long long>((total / 4 + 255) / 256, (long long)sm_count() * 16);
    cseg_launch(embed_tokens_vec4_kernel, dim3(vb), dim3(256), 0, (cudaStream_t)stream, (const float4*)pe, (const float4*)cls,
                                                                  (const float4*)pos, n_crops, L, width / 4, (float4*)x);
  } else {
    cseg_launch(embed_tokens_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, pe, cls, pos, n_crops, L, width, x);
  }
  CSEG_LAUNCH_CHECK("embed_tokens");
  return 0;
}

int cseg_embed_tokens_ln(const float* pe, const float* cls, const float* pos, int n_crops, int L, int width,
                         const float* gamma, const float* beta, float eps, float* x, void* stream) {
  CSEG_REQUIRE(n_crops > 0 && L > 1 && width > 0, "embed_tokens_ln: bad shape");
  CSEG_REQUIRE(width % 4 == 0 && width <= 128 * LNV, "embed_tokens_ln: width=%d must be a multiple of 4 and <= %d", width, 128 * LNV);
  CSEG_REQUIRE((((uintptr_t)pe | (uintptr_t)cls | (uintptr_t)pos | (uintptr_t)x | (uintptr_t)gamma | (uintptr_t)beta) & 15) == 0,
               "embed_tokens_ln: pointers must be 16-byte aligned");
  const long long rows = (long long)n_crops * L;
  cseg_launch(embed_ln_kernel, dim3(cdiv(rows * 32, 256)), dim3(256), 0, (cudaStream_t)stream, (const float4*)pe, (const float4*)cls,
                                                                        (const float4*)pos, n_crops, L, width, (const float4*)gamma,
                                                                        (const float4*)beta, eps, (float4*)x);
  CSEG_LAUNCH_CHECK("embed_tokens_ln");
  return 0;
}

int cseg_layernorm(const float* x, int rows, int width, const float* gamma, const float* beta, float eps,
                   int out_dtype, void* out, void* stream) {
  CSEG_REQUIRE(rows > 0 && width > 0, "layernorm: bad shape");
  const int blocks = cdiv((long long)rows * 32, 256);
  const int pblocks = std::min(blocks, sm_count() * 6);        // register kernel: persistent, grid-stride over rows
  const bool reg_path = (width % 4 == 0) && width <= 128 * LNV && (((uintptr_t)x | (uintptr_t)out | (uintptr_t)gamma | (uintptr_t)beta) & 15) == 0;
  if (reg_path) {
    const int need = cdiv(width / 4, 32);          // float4 vectors per lane
#define CSEG_LN_LAUNCH(NVV)                                                                                                        \
    do {                                                                                                                           \
      if (out_dtype == CSEG_BF16)                                                                                                  \
        cseg_launch(layernorm_reg_kernel<bf16, NVV>, dim3(pblocks), dim3(256), 0, (cudaStream_t)stream, x, rows, width, gamma, beta, eps, (bf16*)out); \
      else                                                                                                                         \
        cseg_launch(layernorm_reg_kernel<float, NVV>, dim3(pblocks), dim3(256), 0, (cudaStream_t)stream, x, rows, width, gamma, beta, eps, (float*)out); \
    } while (0)
    if (need <= 2) CSEG_LN_LAUNCH(2);
    else if (need <= 4) CSEG_LN_LAUNCH(4);
    else if (need <= 6) CSEG_LN_LAUNCH(6);
    else if (need <= 8) CSEG_LN_LAUNCH(8);
    else CSEG_LN_LAUNCH(10);
#undef CSEG_LN_LAUNCH
  } else if (out_dtype == CSEG_BF16)
    cseg_launch(layernorm_kernel<bf16>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, x, rows, width, gamma, beta, eps, (bf16*)out);
  else
    cseg_launch(layernorm_kernel<float>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, x, rows, width, gamma, beta, eps, (float*)out);
  CSEG_LAUNCH_CHECK("layernorm");
  return 0;
}

}  // extern "C"

template <typename T, int HD>
static int launch_attention(const void* qkv, int n_crops, int L, int heads, int mode, const float* simmap,
                            float simw, void* out, float* stats, cudaStream_t st) {
  constexpr int LDS = HD + (sizeof(T) == 2 ? 2 : 1);
  const size_t smem = (((size_t)3 * L * LDS * sizeof(T) + 15) & ~(size_t)15) + (size_t)ATT_WARPS * ATT_JMAX * 32 * 4;
  if (L > ATT_JMAX * 32 || smem > 227 * 1024) {          // long sequences: score rows in shared memory, K / V streamed
    int warps = (int)std::min<size_t>(8, (size_t)(200 * 1024) / ((size_t)L * 4));
    CSEG_REQUIRE(warps >= 1, "attention: L=%d exceeds the %d tokens one score row in shared memory allows", L, 200 * 1024 / 4);
    const size_t sm = (size_t)warps * L * 4;
    const int rows_per_cta = 8 * warps;
    CSEG_SET_SMEM((attention_long_kernel<T, HD>), sm);
    cseg_launch(attention_long_kernel<T, HD>, dim3(n_crops * heads * cdiv(L, rows_per_cta)), dim3(warps * 32), sm, st, (const T*)qkv, L,
                heads, mode, simmap, simw, (T*)out, stats, rows_per_cta);
    CSEG_LAUNCH_CHECK("attention_long");
    return 0;
  }
  CSEG_SET_SMEM((attention_kernel<T, HD>), smem);
  cseg_launch(attention_kernel<T, HD>, dim3(n_crops * heads), dim3(ATT_WARPS * 32), smem, st, (const T*)qkv, L, heads, mode, simmap, simw,
                                                                         (T*)out, stats);
  CSEG_LAUNCH_CHECK("attention");
  return 0;
}

int cseg_attention_tc(const bf16* qkv, int n_crops, int L, int heads, int head_dim, int mode, const float* simmap, float simw,
                      bf16* out, float* stats, cudaStream_t st);
int cseg_gram_split_tc(const void* X, int ldx, int M, int width, int block_rows, int skip, float alpha, float* out,
                       int layout, int cols_pad, cudaStream_t st);

extern "C" {

int cseg_attention(int dtype, const void* qkv, int n_crops, int L, int heads, int head_dim, int mode,
                   const float* simmap, float sim_weight, void* out, float* stats, void* stream) {
  CSEG_REQUIRE(n_crops > 0 && heads > 0, "attention: bad shape");
  CSEG_REQUIRE(L >= 2, "attention: L=%d", L);
  CSEG_REQUIRE(mode >= CSEG_ATTN_STD && mode <= CSEG_ATTN_CAUSAL, "attention: unknown mode %d", mode);
  CSEG_REQUIRE(stats == nullptr || mode == CSEG_ATTN_STD, "attention: stats only with CSEG_ATTN_STD");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == CSEG_BF16) {   // tcgen05 kernel for the standard layers (head_dim 64, L <= 272: ViT-B/16 and ViT-L/14 crops)
    const int rc = cseg_attention_tc((const bf16*)qkv, n_crops, L, heads, head_dim, mode, simmap, sim_weight, (bf16*)out, stats, st);
    if (rc <= 0) return rc;
  }
  if (dtype == CSEG_BF16) {   // mma.sync kernel; returns 1 for shapes it does not cover
    const int rc = cseg_attention_mma((const bf16*)qkv, n_crops, L, heads, head_dim, mode, simmap, sim_weight, (bf16*)out,
                                      stats, st);
    if (rc <= 0) return rc;
  }
  if (dtype == CSEG_BF16 && head_dim == 64)
    return launch_attention<bf16, 64>(qkv, n_crops, L, heads, mode, simmap, sim_weight, out, stats, st);
  if (dtype == CSEG_F32 && head_dim == 64)
    return launch_attention<float, 64>(qkv, n_crops, L, heads, mode, simmap, sim_weight, out, stats, st);
  if (dtype == CSEG_BF16 && head_dim == 32)
    return launch_attention<bf16, 32>(qkv, n_crops, L, heads, mode, simmap, sim_weight, out, stats, st);
  if (dtype == CSEG_F32 && head_dim == 32)
    return launch_attention<float, 32>(qkv, n_crops, L, heads, mode, simmap, sim_weight, out, stats, st);
  if (dtype == CSEG_BF16 && head_dim == 80)
    return launch_attention<bf16, 80>(qkv, n_crops, L, heads, mode, simmap, sim_weight, out, stats, st);
  if (dtype == CSEG_F32 && head_dim == 80)
    return launch_attention<float, 80>(qkv, n_crops, L, heads, mode, simmap, sim_weight, out, stats, st);
  CSEG_FAIL(CSEG_EUNSUPPORTED, "attention: head_dim=%d dtype=%d not supported (32, 64 or 80)", head_dim, dtype);
}

int cseg_simmap(const float* x, int n_crops, int L, int width, float temperature, int add_self_similarity,
                float* simmap, void* stream) {
  CSEG_REQUIRE(n_crops > 0 && L >= 2 && width > 0 && temperature != 0.f, "simmap: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int P = L - 1;
  dim3 grid(cdiv(P, 64), cdiv(P, 64), n_crops);
  cseg_launch(simmap_kernel, dim3(grid), dim3(256), 0, st, x, L, width, 1.0f / temperature, add_self_similarity, simmap);
  CSEG_LAUNCH_CHECK("simmap");
  return 0;
}

int cseg_simmap_tc(const float* x, int n_crops, int L, int width, float temperature, void* scratch, float* simmap,
                   int layout, void* stream) {
  CSEG_REQUIRE(layout == 0 || (layout == 1 && L <= CSEG_SIMT_COLS_MAX), "simmap_tc: layout=%d L=%d", layout, L);
  CSEG_REQUIRE(n_crops > 0 && L >= 2 && width > 0 && temperature != 0.f, "simmap_tc: bad arguments");
  CSEG_REQUIRE(width % 64 == 0 && width <= 128 * LNV, "simmap_tc: width=%d must be a multiple of 64 and <= %d", width, 128 * LNV);
  CSEG_REQUIRE((((uintptr_t)x | (uintptr_t)scratch) & 15) == 0, "simmap_tc: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const long long rows = (long long)n_crops * L;
  cseg_launch(simmap_split_kernel, dim3(cdiv(rows * 32, 256)), dim3(256), 0, st, x, rows, width, (bf16*)scratch);
  CSEG_LAUNCH_CHECK("simmap_split");
  return cseg_gram_split_tc(scratch, 2 * width, (int)rows, width, L, 1, 1.0f / temperature, simmap, layout, CSEG_SIMT_COLS_FOR(L), st);
}

int cseg_outlier_suppress(const float* y, float* y_out, int n_crops, int L, int width, int grid, const float* stats,
                          int heads, int top_k, float contamination_temp, int32_t* plan, int32_t* outlier_idx,
                          void* stream) {
  CSEG_REQUIRE(n_crops > 0 && grid * grid == L - 1, "outlier_suppress: grid %d^2 != L-1 = %d", grid, L - 1);
  CSEG_REQUIRE(top_k > 0 && top_k <= OS_MAXK, "outlier_suppress: top_k=%d outside [1, %d]", top_k, OS_MAXK);
  CSEG_REQUIRE(y != y_out, "outlier_suppress: runs out of place (y_out must differ from y)");
  const size_t smem = (size_t)(L - 1) * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  cseg_launch(outlier_plan_kernel, dim3(n_crops), dim3(OS_THREADS), smem, st, y, L, width, grid, stats, heads, top_k, plan, outlier_idx);
  CSEG_LAUNCH_CHECK("outlier_plan");
  cseg_launch(outlier_apply_kernel, dim3(n_crops * L), dim3(128), 0, st, y, y_out, L, width, top_k, contamination_temp, plan);
  CSEG_LAUNCH_CHECK("outlier_apply");
  return 0;
}

int cseg_cls_debias(const float* tok, int n_crops, int L, int D, float factor, int out_dtype, void* feats, int ldf,
                    int rows_per_crop, float* cls_unit, void* stream) {
  CSEG_REQUIRE(n_crops > 0 && L >= 2 && D > 0 && ldf >= D, "cls_debias: bad shape");
  const int rows = rows_per_crop > 0 ? rows_per_crop : L - 1;
  CSEG_REQUIRE(rows >= L - 1, "cls_debias: rows_per_crop=%d < %d patch tokens", rows, L - 1);
  const int blocks = cdiv((long long)n_crops * (rows + 1) * 32, 256);
  if (out_dtype == CSEG_BF16)
    cseg_launch(cls_debias_kernel<bf16>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, tok, n_crops, L, D, factor, (bf16*)feats, ldf,
                                                                      rows, cls_unit);
  else
    cseg_launch(cls_debias_kernel<float>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, tok, n_crops, L, D, factor, (float*)feats, ldf,
                                                                       rows, cls_unit);
  CSEG_LAUNCH_CHECK("cls_debias");
  return 0;
}

}  // extern "C"
