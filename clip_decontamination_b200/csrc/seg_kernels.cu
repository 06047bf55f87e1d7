// Segmentor-side kernels: per-pixel L2-normalise + cosine logits (segmentor.py:374-375), the fused
// sliding-window accumulate -> count-normalise -> resize -> softmax -> synonym max -> argmax ->
// background threshold (segmentor.py:413-449,475-489) and the IoU histograms of mmseg's IoUMetric.
#include "common.cuh"

namespace {

constexpr int QMAX = 32;  // queries per class file (largest shipped: iSAID 16, OpenEarthMap 8+)

// ---------------------------------------------------------------------------------------------
// A10: logits[crop][q][pix] = <f/|f|, t_q> (+ cls bias).  8 lanes per feature row, 16-byte loads.
// HBM-bound: reads rows*D*sizeof(T), writes rows*Q*4.
// ---------------------------------------------------------------------------------------------
template <typename T> struct Vec8;
template <> struct Vec8<bf16> {
  static __device__ __forceinline__ void ld(const bf16* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
  }
};
template <> struct Vec8<float> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};

template <typename T, int QT>
__global__ void __launch_bounds__(256) norm_sim_kernel(const T* __restrict__ feats, int ldf, long long rows, int hw,
                                                       int D, const float* __restrict__ text, int Q,
                                                       const float* __restrict__ cls_bias,
                                                       float* __restrict__ logits) {
  pdl_grid_sync();
  extern __shared__ float ts[];  // [Q][D]
  for (int i = threadIdx.x; i < Q * D; i += blockDim.x) ts[i] = text[i];
  __syncthreads();
  const int sub = threadIdx.x & 7;
  for (long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3; row < rows;
       row += ((long long)gridDim.x * blockDim.x) >> 3) {
    const T* f = feats + row * ldf;
    float ss = 0.f, dot[QT];
#pragma unroll
    for (int q = 0; q < QT; ++q) dot[q] = 0.f;
    for (int c = sub * 8; c < D; c += 64) {
      float v[8];
      Vec8<T>::ld(f + c, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) ss = fmaf(v[e], v[e], ss);
#pragma unroll
      for (int q = 0; q < QT; ++q)
        if (q < Q) {
          const float* t = ts + q * D + c;
#pragma unroll
          for (int e = 0; e < 8; ++e) dot[q] = fmaf(v[e], t[e], dot[q]);
        }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      ss += __shfl_xor_sync(0xffffffffu, ss, o);
#pragma unroll
      for (int q = 0; q < QT; ++q) dot[q] += __shfl_xor_sync(0xffffffffu, dot[q], o);
    }
    const float inv = 1.0f / sqrtf(ss);
    const long long crop = row / hw, pix = row % hw;
#pragma unroll
    for (int q = 0; q < QT; ++q)
      if (q < Q && (q & 7) == sub) {
        float v = dot[q] * inv;
        if (cls_bias) v += cls_bias[crop * Q + q];
        logits[(crop * Q + q) * hw + pix] = v;
      }
  }
}

// ---------------------------------------------------------------------------------------------
// A11 + A12 fused.  Every crop logit is read exactly once when (out_h,out_w) == (H,W); labels are the only
// mandatory write.
// ---------------------------------------------------------------------------------------------
struct AccumParams {
  const float* crop_logits;
  int n_crops, Q, lh, lw, crop_h, crop_w, pad_top, pad_left;
  const int32_t* windows;
  int H, W, out_h, out_w;
  const int32_t* query_idx;
  int K;
  float logit_scale, prob_thd;
  int bg_idx;
  uint8_t* labels;
  float* probs;
  float* avg_logits;
};

// torch upsample_bilinear2d, align_corners=False: src = max(scale*(dst+0.5)-0.5, 0)
__device__ __forceinline__ void bilin_coord(int dst, int in_size, int out_size, int& i0, int& i1, float& l1) {
  const float scale = (float)in_size / (float)out_size;
  float src = scale * ((float)dst + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  i0 = (int)src;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = src - (float)i0;
}

// value of crop `cr`, query q at canvas-crop coordinate (cy, cx) in [0,crop_h) x [0,crop_w)
__device__ __forceinline__ float crop_value(const AccumParams& p, int cr, int q, int cy, int cx) {
  const float* base = p.crop_logits + ((size_t)cr * p.Q + q) * p.lh * p.lw;
  if (p.lh == p.crop_h && p.lw == p.crop_w) return base[cy * p.lw + cx];
  int y0, y1, x0, x1;
  float ly, lx;
  bilin_coord(cy, p.lh, p.crop_h, y0, y1, ly);
  bilin_coord(cx, p.lw, p.crop_w, x0, x1, lx);
  const float hy = 1.f - ly, hx = 1.f - lx;
  return hy * (hx * base[y0 * p.lw + x0] + lx * base[y0 * p.lw + x1]) +
         ly * (hx * base[y1 * p.lw + x0] + lx * base[y1 * p.lw + x1]);
}

// averaged logits of canvas pixel (y, x): candidate windows visited in forward_slide order (segmentor.py:416-444)
template <int QT>
__device__ __forceinline__ void canvas_avg(const AccumParams& p, const int4* wins, const int* cand, int n_cand, int y,
                                           int x, float (&acc)[QT]) {
#pragma unroll
  for (int q = 0; q < QT; ++q) acc[q] = 0.f;
  int count = 0;
  for (int ci = 0; ci < n_cand; ++ci) {
    const int cr = cand[ci];
    const int4 w = wins[ci];  // y1, x1, h, w
    const int ly = y - w.x, lx = x - w.y;
    if (ly < 0 || ly >= w.z || lx < 0 || lx >= w.w) continue;
    ++count;
#pragma unroll
    for (int q = 0; q < QT; ++q)
      if (q < p.Q) acc[q] += crop_value(p, cr, q, ly + p.pad_top, lx + p.pad_left);
  }
  const float cnt = (float)count;
#pragma unroll
  for (int q = 0; q < QT; ++q)
    if (q < p.Q) acc[q] = acc[q] / cnt;
}

// x logit_scale, softmax over Q, per-class max over synonym queries, argmax (lowest index wins ties), threshold
// (segmentor.py:478-489).  Ordering is decided on the scaled logits (softmax is monotone); probabilities only feed the
// threshold and the optional probs output.
template <int QT>
__device__ __forceinline__ uint8_t post_process(const AccumParams& p, const int* s_qidx, bool identity, float (&v)[QT], int oy,
                                                int ox) {
  float m = -INFINITY;
  int arg = 0;
#pragma unroll
  for (int q = 0; q < QT; ++q)
    if (q < p.Q) {
      v[q] *= p.logit_scale;
      if (v[q] > m) {          // strict: the lowest index wins ties
        m = v[q];
        arg = q;
      }
    }
  float sum = 0.f;
#pragma unroll
  for (int q = 0; q < QT; ++q)
    if (q < p.Q) sum += expf(v[q] - m);
  if (identity && p.probs == nullptr) {
    // one query per class (query_idx = 0..Q-1): the class maximum is the query maximum, its probability 1 / sum
    return (uint8_t)((1.0f / sum < p.prob_thd) ? p.bg_idx : arg);
  }
  float best = -INFINITY;
  int best_k = 0;
  for (int k = 0; k < p.K; ++k) {
    float ck = -INFINITY;
#pragma unroll
    for (int q = 0; q < QT; ++q)
      if (q < p.Q && s_qidx[q] == k) ck = fmaxf(ck, v[q]);
    if (p.probs) p.probs[((size_t)k * p.out_h + oy) * p.out_w + ox] = expf(ck - m) / sum;
    if (ck > best) {
      best = ck;
      best_k = k;
    }
  }
  const float pmax = expf(best - m) / sum;
  if (pmax < p.prob_thd) best_k = p.bg_idx;  // segmentor.py:489
  return (uint8_t)best_k;
}

// One CTA per tile of (32 PX) x 8 output pixels, PX consecutive pixels per thread.  The windows that touch the tile are
// binned once per CTA (in forward_slide order, so the fp32 sums keep the reference's order): a pixel then visits
// <= 4..9 candidates instead of all n_crops.  When the crop logits are at crop resolution and no final resize is needed
// (the JBU path) the PX pixels of a thread read each crop's logits with one 16-byte load per query.
template <int QT, int PX, bool DIRECT>
__global__ void __launch_bounds__(256, 2) accum_argmax_kernel(const AccumParams p) {
  pdl_grid_sync();
  extern __shared__ int4 s_wins[];                              // [n_crops] candidate windows (compacted)
  int* s_cand = reinterpret_cast<int*>(s_wins + p.n_crops);     // [n_crops] their crop indices
  __shared__ int s_ncand;
  __shared__ int s_qidx[QT];
  __shared__ int s_ident;
  const int tid = threadIdx.x;
  const int ox0 = blockIdx.x * (32 * PX), oy0 = blockIdx.y * 8;
  constexpr bool direct = DIRECT;                              // output size == canvas size (no final resize)
  if (tid < QT) s_qidx[tid] = tid < p.Q ? p.query_idx[tid] : -1;
  if (tid == 32) {
    int id = (p.K == p.Q);
    for (int q = 0; q < p.Q; ++q) id &= (p.query_idx[q] == q);
    s_ident = id;
  }
  if (tid < 32) {
    // canvas rectangle the tile depends on
    int cy0 = oy0, cy1 = min(oy0 + 7, p.out_h - 1), cx0 = ox0, cx1 = min(ox0 + 32 * PX - 1, p.out_w - 1);
    if (!direct) {
      int a, b;
      float l;
      bilin_coord(cy0, p.H, p.out_h, a, b, l); cy0 = a;
      bilin_coord(cy1, p.H, p.out_h, a, b, l); cy1 = b;
      bilin_coord(cx0, p.W, p.out_w, a, b, l); cx0 = a;
      bilin_coord(cx1, p.W, p.out_w, a, b, l); cx1 = b;
    }
    int n = 0;
    for (int base = 0; base < p.n_crops; base += 32) {
      const int i = base + tid;
      int4 w = make_int4(0, 0, 0, 0);
      bool hit = false;
      if (i < p.n_crops) {
        w = __ldg(reinterpret_cast<const int4*>(p.windows) + i);
        hit = (w.x <= cy1 && w.x + w.z > cy0 && w.y <= cx1 && w.y + w.w > cx0);
      }
      const unsigned mask = __ballot_sync(0xffffffffu, hit);
      if (hit) {
        const int pos = n + __popc(mask & ((1u << tid) - 1u));
        s_wins[pos] = w;
        s_cand[pos] = i;
      }
      n += __popc(mask);
    }
    if (tid == 0) s_ncand = n;
  }
  __syncthreads();
  const int n_cand = s_ncand;
  const bool identity = s_ident != 0;
  const int oy = oy0 + (tid >> 5), oxb = ox0 + (tid & 31) * PX;
  if (oy >= p.out_h || oxb >= p.out_w) return;
  uint8_t lab[PX];
  if (direct) {
    float acc[PX][QT];
    int cnt[PX];
#pragma unroll
    for (int e = 0; e < PX; ++e) {
      cnt[e] = 0;
#pragma unroll
      for (int q = 0; q < QT; ++q) acc[e][q] = 0.f;
    }
    const bool same_res = (p.lh == p.crop_h && p.lw == p.crop_w);
    // candidate that covers all PX pixels of this thread with one aligned vector per query: its base pointer (else null)
    auto vec_base = [&](int ci) -> const float* {
      const int4 w = s_wins[ci];
      const int ly = oy - w.x, lx0 = oxb - w.y;
      if (ly < 0 || ly >= w.z || lx0 < 0 || lx0 + PX > w.w) return nullptr;
      const int cy = ly + p.pad_top, cx0 = lx0 + p.pad_left;
      if (((cx0 | p.lw) & (PX - 1)) != 0) return nullptr;
      return p.crop_logits + (size_t)s_cand[ci] * p.Q * p.lh * p.lw + (size_t)cy * p.lw + cx0;
    };
    auto vec_load = [&](const float* base, float (&v)[QT][PX]) {
#pragma unroll
      for (int q = 0; q < QT; ++q)
        if (q < p.Q) {
          if (PX == 4) {
            const float4 v4 = *reinterpret_cast<const float4*>(base + (size_t)q * p.lh * p.lw);
            v[q][0] = v4.x; v[q][1 % PX] = v4.y; v[q][2 % PX] = v4.z; v[q][3 % PX] = v4.w;
          } else {
            const float2 v2 = *reinterpret_cast<const float2*>(base + (size_t)q * p.lh * p.lw);
            v[q][0] = v2.x; v[q][1 % PX] = v2.y;
          }
        }
    };
    auto vec_add = [&](const float (&v)[QT][PX]) {
#pragma unroll
      for (int q = 0; q < QT; ++q)
        if (q < p.Q) {
#pragma unroll
          for (int e = 0; e < PX; ++e) acc[e][q] += v[q][e];
        }
#pragma unroll
      for (int e = 0; e < PX; ++e) ++cnt[e];
    };
    for (int ci = 0; ci < n_cand; ++ci) {
      if (PX >= 2 && same_res) {
        // the loads of TWO consecutive vector candidates are in flight together (the kernel was latency bound: one
        // candidate's 16-byte loads, a DRAM round trip, the next candidate's ...); the sums keep forward_slide's order
        const float* b0 = vec_base(ci);
        if (b0 != nullptr) {
          const float* b1 = (QT <= 8 && ci + 1 < n_cand) ? vec_base(ci + 1) : nullptr;    // (register budget: Q <= 8 only)
          float v0[QT][PX], v1[QT][PX];
          vec_load(b0, v0);
          if (b1 != nullptr) vec_load(b1, v1);
          vec_add(v0);
          if (b1 != nullptr) {
            vec_add(v1);
            ++ci;
          }
          continue;
        }
      }
      const int4 w = s_wins[ci];
      const int ly = oy - w.x;
      if (ly < 0 || ly >= w.z) continue;
      const int lx0 = oxb - w.y;
      if (lx0 + PX <= 0 || lx0 >= w.w) continue;
      const int cr = s_cand[ci];
      const int cy = ly + p.pad_top, cx0 = lx0 + p.pad_left;
#pragma unroll
      for (int e = 0; e < PX; ++e) {
        const int lx = lx0 + e;
        if (lx < 0 || lx >= w.w || oxb + e >= p.out_w) continue;
        ++cnt[e];
#pragma unroll
        for (int q = 0; q < QT; ++q)
          if (q < p.Q) acc[e][q] += crop_value(p, cr, q, cy, cx0 + e);
      }
    }
#pragma unroll
    for (int e = 0; e < PX; ++e) {
      const int ox = oxb + e;
      if (ox >= p.out_w) { lab[e] = 0; continue; }
      const float c = (float)cnt[e];
#pragma unroll
      for (int q = 0; q < QT; ++q)
        if (q < p.Q) {
          acc[e][q] = acc[e][q] / c;
          if (p.avg_logits) p.avg_logits[((size_t)q * p.H + oy) * p.W + ox] = acc[e][q];
        }
      lab[e] = post_process<QT>(p, s_qidx, identity, acc[e], oy, ox);
    }
  } else {  // bilinear resize of the averaged canvas to ori_shape (segmentor.py:448-449)
    int y0, y1;
    float ly;
    bilin_coord(oy, p.H, p.out_h, y0, y1, ly);
    const float hy = 1.f - ly;
#pragma unroll 1
    for (int e = 0; e < PX; ++e) {
      const int ox = oxb + e;
      if (ox >= p.out_w) { lab[e] = 0; continue; }
      int x0, x1;
      float lx;
      bilin_coord(ox, p.W, p.out_w, x0, x1, lx);
      const float hx = 1.f - lx;
      float v[QT], a[QT], b[QT];
      canvas_avg<QT>(p, s_wins, s_cand, n_cand, y0, x0, a);
      canvas_avg<QT>(p, s_wins, s_cand, n_cand, y0, x1, b);
#pragma unroll
      for (int q = 0; q < QT; ++q) v[q] = hy * (hx * a[q] + lx * b[q]);
      canvas_avg<QT>(p, s_wins, s_cand, n_cand, y1, x0, a);
      canvas_avg<QT>(p, s_wins, s_cand, n_cand, y1, x1, b);
#pragma unroll
      for (int q = 0; q < QT; ++q) v[q] += ly * (hx * a[q] + lx * b[q]);
      lab[e] = post_process<QT>(p, s_qidx, identity, v, oy, ox);
    }
  }
  uint8_t* lrow = p.labels + (size_t)oy * p.out_w + oxb;
  if (PX == 4 && oxb + 4 <= p.out_w && (((size_t)oy * p.out_w + oxb) & 3) == 0 && ((uintptr_t)p.labels & 3) == 0) {
    *reinterpret_cast<uint32_t*>(lrow) = (uint32_t)lab[0] | ((uint32_t)lab[1 % PX] << 8) | ((uint32_t)lab[2 % PX] << 16) |
                                         ((uint32_t)lab[3 % PX] << 24);
  } else if (PX == 2 && oxb + 2 <= p.out_w && (((size_t)oy * p.out_w + oxb) & 1) == 0 && ((uintptr_t)p.labels & 1) == 0) {
    *reinterpret_cast<unsigned short*>(lrow) = (unsigned short)((uint32_t)lab[0] | ((uint32_t)lab[1 % PX] << 8));
  } else {
#pragma unroll
    for (int e = 0; e < PX; ++e)
      if (oxb + e < p.out_w) lrow[e] = lab[e];
  }
}

// ---------------------------------------------------------------------------------------------
// K17: per-class intersect / pred / label counts (mmseg IoUMetric.intersect_and_union)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) iou_hist_kernel(const uint8_t* __restrict__ pred,
                                                       const uint8_t* __restrict__ label, long long n, int K,
                                                       int ignore, unsigned long long* __restrict__ hist) {
  pdl_grid_sync();
  extern __shared__ unsigned int sh[];  // [3][K]
  for (int i = threadIdx.x; i < 3 * K; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int l = label[i], pr = pred[i];
    if (l == ignore) continue;
    if (pr < K) atomicAdd(&sh[K + pr], 1u);
    if (l < K) atomicAdd(&sh[2 * K + l], 1u);
    if (pr == l && pr < K) atomicAdd(&sh[pr], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * K; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

// ---------------------------------------------------------------------------------------------
// N4 output side: colourised mask and confidence heat-map (segmentor.py:513-531,568-608) as BGR images
// ready for cv2.imwrite.  lut: uint8 [n_lut][3] (palette rows for the mask; the 256-entry colour map for the heat-map).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colorize_kernel(const uint8_t* __restrict__ labels, long long n,
                                                       const uint8_t* __restrict__ lut, int n_lut,
                                                       uint8_t* __restrict__ out) {
  pdl_grid_sync();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int l = min((int)labels[i], n_lut - 1);          // np.clip(mask, 0, len(palette) - 1)
    out[3 * i + 0] = lut[3 * l + 0];
    out[3 * i + 1] = lut[3 * l + 1];
    out[3 * i + 2] = lut[3 * l + 2];
  }
}
__global__ void __launch_bounds__(256) heatmap_kernel(const float* __restrict__ probs, int K, long long n,
                                                      const uint8_t* __restrict__ lut, uint8_t* __restrict__ out) {
  pdl_grid_sync();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float c = -INFINITY;
    for (int k = 0; k < K; ++k) c = fmaxf(c, probs[(size_t)k * n + i]);     // seg_logits.max(dim=0)
    if (!(c == c)) c = 0.f;                                                  // nan_to_num
    c = fminf(fmaxf(c, 0.f), 1.f);
    const int g = (int)(c * 255.0f);                                         // astype(uint8) truncates
    out[3 * i + 0] = lut[3 * g + 0];
    out[3 * i + 1] = lut[3 * g + 1];
    out[3 * i + 2] = lut[3 * g + 2];
  }
}

}  // namespace

template <typename T>
static int launch_norm_sim(const void* feats, int ldf, int n_crops, int hw, int D, const float* text, int Q,
                           const float* cls_bias, float* logits, cudaStream_t st) {
  const long long rows = (long long)n_crops * hw;
  const size_t smem = (size_t)Q * D * sizeof(float);
  const int blocks = (int)std::min<long long>((rows * 8 + 255) / 256, (long long)sm_count() * 8);
#define NS_LAUNCH(QT)                                                                                          \
  do {                                                                                                         \
    CSEG_SET_SMEM((norm_sim_kernel<T, QT>), smem);                                                              \
    cseg_launch(norm_sim_kernel<T, QT>, dim3(blocks), dim3(256), smem, st, (const T*)feats, ldf, rows, hw, D, text, Q, cls_bias, logits); \
  } while (0)
  if (Q <= 8) NS_LAUNCH(8);
  else if (Q <= 16) NS_LAUNCH(16);
  else NS_LAUNCH(32);
#undef NS_LAUNCH
  CSEG_LAUNCH_CHECK("norm_sim");
  return 0;
}

extern "C" {

int cseg_norm_sim(int dtype, const void* feats, int ldf, int n_crops, int hw, int D, const float* text, int Q,
                  const float* cls_logit_bias, float* logits, void* stream) {
  CSEG_REQUIRE(n_crops > 0 && hw > 0 && D > 0 && D % 8 == 0 && ldf % 8 == 0, "norm_sim: D=%d, ldf=%d must be multiples of 8", D, ldf);
  CSEG_REQUIRE(Q >= 1 && Q <= QMAX, "norm_sim: Q=%d outside [1, %d]", Q, QMAX);
  CSEG_REQUIRE((size_t)Q * D * 4 <= 200 * 1024, "norm_sim: text table too large for shared memory");
  if (dtype == CSEG_BF16)
    return launch_norm_sim<bf16>(feats, ldf, n_crops, hw, D, text, Q, cls_logit_bias, logits, (cudaStream_t)stream);
  return launch_norm_sim<float>(feats, ldf, n_crops, hw, D, text, Q, cls_logit_bias, logits, (cudaStream_t)stream);
}

int cseg_accum_argmax(const float* crop_logits, int n_crops, int Q, int lh, int lw, int crop_h, int crop_w,
                      int pad_top, int pad_left, const int32_t* windows, int H, int W, int out_h, int out_w,
                      const int32_t* query_idx, int K, float logit_scale, float prob_thd, int bg_idx,
                      uint8_t* labels, float* probs, float* avg_logits, void* stream) {
  CSEG_REQUIRE(n_crops > 0 && H > 0 && W > 0 && out_h > 0 && out_w > 0, "accum_argmax: empty problem");
  CSEG_REQUIRE(Q >= 1 && Q <= QMAX, "accum_argmax: Q=%d outside [1, %d]", Q, QMAX);
  CSEG_REQUIRE(K >= 1 && K <= 256 && bg_idx >= 0 && bg_idx < 256, "accum_argmax: K=%d / bg_idx=%d do not fit uint8 labels", K, bg_idx);
  CSEG_REQUIRE(avg_logits == nullptr || (out_h == H && out_w == W), "accum_argmax: avg_logits needs out size == canvas size");
  AccumParams p{crop_logits, n_crops, Q, lh, lw, crop_h, crop_w, pad_top, pad_left, windows, H, W, out_h, out_w,
                query_idx, K, logit_scale, prob_thd, bg_idx, labels, probs, avg_logits};
  const size_t smem = (size_t)n_crops * (sizeof(int4) + sizeof(int));
  CSEG_REQUIRE(smem <= 200 * 1024, "accum_argmax: %d windows do not fit the window table in shared memory", n_crops);
  const bool direct = (out_h == H && out_w == W);
#define CSEG_ACCUM_LAUNCH(QT, PX, D)                                                                                  \
  do {                                                                                                                \
    CSEG_SET_SMEM((accum_argmax_kernel<QT, PX, D>), smem);                                                            \
    cseg_launch(accum_argmax_kernel<QT, PX, D>, dim3(cdiv(out_w, 32 * PX), cdiv(out_h, 8)), dim3(256), smem,          \
                (cudaStream_t)stream, p);                                                                             \
  } while (0)
  if (Q <= 8) {
    if (direct) CSEG_ACCUM_LAUNCH(8, 4, true); else CSEG_ACCUM_LAUNCH(8, 4, false);
  } else if (Q <= 16) {
    if (direct) CSEG_ACCUM_LAUNCH(16, 2, true); else CSEG_ACCUM_LAUNCH(16, 2, false);
  } else {
    if (direct) CSEG_ACCUM_LAUNCH(32, 1, true); else CSEG_ACCUM_LAUNCH(32, 1, false);
  }
#undef CSEG_ACCUM_LAUNCH
  CSEG_LAUNCH_CHECK("accum_argmax");
  return 0;
}

int cseg_colorize(const uint8_t* labels, long long n, const uint8_t* lut, int n_lut, uint8_t* out_bgr, void* stream) {
  CSEG_REQUIRE(labels && lut && out_bgr && n > 0 && n_lut > 0, "colorize: bad arguments");
  const int blocks = (int)std::min<long long>((n + 255) / 256, (long long)sm_count() * 16);
  cseg_launch(colorize_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, labels, n, lut, n_lut, out_bgr);
  CSEG_LAUNCH_CHECK("colorize");
  return 0;
}

int cseg_heatmap(const float* probs, int K, long long n, const uint8_t* lut256, uint8_t* out_bgr, void* stream) {
  CSEG_REQUIRE(probs && lut256 && out_bgr && n > 0 && K > 0, "heatmap: bad arguments");
  const int blocks = (int)std::min<long long>((n + 255) / 256, (long long)sm_count() * 16);
  cseg_launch(heatmap_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, probs, K, n, lut256, out_bgr);
  CSEG_LAUNCH_CHECK("heatmap");
  return 0;
}

int cseg_iou_hist(const uint8_t* pred, const uint8_t* label, long long n, int K, int ignore_index, long long* hist,
                  void* stream) {
  CSEG_REQUIRE(n > 0 && K >= 1 && K <= 256, "iou_hist: bad arguments");
  const int blocks = (int)std::min<long long>((n + 255) / 256, (long long)sm_count() * 8);
  cseg_launch(iou_hist_kernel, dim3(blocks), dim3(256), (size_t)3 * K * sizeof(unsigned int), (cudaStream_t)stream,
              pred, label, n, K, ignore_index, (unsigned long long*)hist);
  CSEG_LAUNCH_CHECK("iou_hist");
  return 0;
}

}  // extern "C"
