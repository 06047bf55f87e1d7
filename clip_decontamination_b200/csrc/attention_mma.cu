// Tensor-core attention for the bf16 path (head_dim 64, L <= 272): one CTA per (crop, head), Q/K/V of the
// head staged in shared memory, each warp owns 16-query-row blocks.  S = X.Y^T is accumulated with
// mma.sync m16n8k16 (bf16 in, fp32 accumulate) for ALL keys at once (L <= 272 keys fit in registers, so no
// online softmax is needed); the softmaxes run in fp32 on the accumulator fragments; P is re-packed to bf16
// A-fragments in registers and O = P.V is a second MMA.  Every custom_attn variant of the reference
// (open_clip/transformer.py:858-908) is one to three such passes:
//   STD / vanilla : softmax(q k^T s [+M]) v                       ClearCLIP : softmax(q q^T s + M) v
//   SFP           : softmax(.5 (qq+kk) s + M) v                   SCLIP     : [softmax(qq s+M) + softmax(kk s+M)] v
//   Experimental  : softmax(softmax((kk+qq) s) + M) v  -- qq and kk accumulate into the SAME fragment
//   SegEarth      : SCLIP + softmax(v v^T s + M) v
// The block-(layers-2) statistics P[0,1+i], P[1+i,1+i] (outlier_suppression.py:46-53) are emitted from the
// probability fragments.
#include "common.cuh"

namespace {

constexpr int HD = 64;
constexpr int RSTRIDE = HD * 2 + 16;  // bytes per staged row (padded: conflict-free ldmatrix)
constexpr int AWARPS = 4;            // general modes: one 16-row block per warp, ceil(L/64) CTAs per (crop, head)
constexpr int SWARPS = 8;            // STD layers: ONE CTA of 8 warps per (crop, head), Q/K/V staged once

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// S[16 rows, NKB*8 keys] += X[rows r0..r0+15, :] . Y[all keys, :]^T   (both staged [row][64] bf16)
template <int NKB>
__device__ __forceinline__ void accum_scores(float (&S)[NKB][4], uint32_t xt, uint32_t yt, int r0, int lane) {
  uint32_t a[4][4];
  {  // A fragments: matrices (rows 0-7,d 0-7) (rows 8-15,d 0-7) (rows 0-7,d 8-15) (rows 8-15,d 8-15)
    const int q = lane >> 3, rr = lane & 7;
    const uint32_t base = xt + (uint32_t)((r0 + (q & 1) * 8 + rr) * RSTRIDE + (q >> 1) * 16);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) ldsm_x4(base + ks * 32, a[ks]);
  }
  // B fragments: keys n (8 per block) x dims; x4 = dims 0-31 of one key block
  const int q = lane >> 3, rr = lane & 7;
#pragma unroll
  for (int nb = 0; nb < NKB; ++nb) {
    uint32_t b[2][4];
    const uint32_t base = yt + (uint32_t)((nb * 8 + rr) * RSTRIDE + q * 16);
    ldsm_x4(base, b[0]);
    ldsm_x4(base + 64, b[1]);
    mma16816(S[nb], a[0], b[0][0], b[0][1]);
    mma16816(S[nb], a[1], b[0][2], b[0][3]);
    mma16816(S[nb], a[2], b[1][0], b[1][1]);
    mma16816(S[nb], a[3], b[1][2], b[1][3]);
  }
}

// row-wise softmax over the valid keys of the fragment rows (row g: elems 0,1; row g+8: elems 2,3)
template <int NKB>
__device__ __forceinline__ void frag_softmax(float (&S)[NKB][4], int L, int tig) {
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int nb = 0; nb < NKB; ++nb) {
    const int c = nb * 8 + 2 * tig;
    if (c >= L) { S[nb][0] = -INFINITY; S[nb][2] = -INFINITY; }
    if (c + 1 >= L) { S[nb][1] = -INFINITY; S[nb][3] = -INFINITY; }
    m0 = fmaxf(m0, fmaxf(S[nb][0], S[nb][1]));
    m1 = fmaxf(m1, fmaxf(S[nb][2], S[nb][3]));
  }
  m0 = quad_max(m0);
  m1 = quad_max(m1);
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int nb = 0; nb < NKB; ++nb) {
    S[nb][0] = __expf(S[nb][0] - m0);
    S[nb][1] = __expf(S[nb][1] - m0);
    S[nb][2] = __expf(S[nb][2] - m1);
    S[nb][3] = __expf(S[nb][3] - m1);
    s0 += S[nb][0] + S[nb][1];
    s1 += S[nb][2] + S[nb][3];
  }
  const float i0 = 1.0f / quad_sum(s0), i1 = 1.0f / quad_sum(s1);
#pragma unroll
  for (int nb = 0; nb < NKB; ++nb) {
    S[nb][0] *= i0; S[nb][1] *= i0; S[nb][2] *= i1; S[nb][3] *= i1;
  }
}

// S += simw * M_pad  (zero CLS row / column, similarity_enhancement.py:104-122)
template <int NKB>
__device__ __forceinline__ void add_simmap(float (&S)[NKB][4], const float* __restrict__ sim, float simw, int P, int L,
                                           int row0, int row1, int tig) {
  if (sim == nullptr) return;
  const float* m0 = (row0 >= 1 && row0 < L) ? sim + (size_t)(row0 - 1) * P - 1 : nullptr;
  const float* m1 = (row1 >= 1 && row1 < L) ? sim + (size_t)(row1 - 1) * P - 1 : nullptr;
#pragma unroll
  for (int nb = 0; nb < NKB; ++nb) {
    const int c = nb * 8 + 2 * tig;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int j = c + e;
      if (j >= 1 && j < L) {
        if (m0) S[nb][e] += simw * __ldg(m0 + j);
        if (m1) S[nb][2 + e] += simw * __ldg(m1 + j);
      }
    }
  }
}

// STD = true: compile-time plain softmax(q k^T s) v (11 of the 12 ViT-B layers): no mode branches, no similarity
// map, one pass -- and one CTA of NW = 8 warps per (crop, head), so Q/K/V are staged once instead of once per
// 64-row split.  STD = false: every custom_attn variant, NW = 4 warps per 64-row split.
template <int NKB, int NW, bool STD>
__global__ void __launch_bounds__(NW * 32, (NW * NKB > 8 * 26) ? 1 : 2) attention_mma_kernel(const bf16* __restrict__ qkv, int L, int heads,
                                                                   int mode_rt, int nsplit, const float* __restrict__ simmap,
                                                                   float simw, bf16* __restrict__ out,
                                                                   float* __restrict__ stats) {
  pdl_grid_sync();
  const int mode = STD ? (int)CSEG_ATTN_STD : mode_rt;
  constexpr int AWARPS = NW;
  constexpr int LP = NKB * 8;  // padded key count (multiple of 16)
  extern __shared__ __align__(16) uint8_t asmem[];
  const uint32_t qt = (uint32_t)__cvta_generic_to_shared(asmem);
  const uint32_t kt = qt + LP * RSTRIDE, vt = kt + LP * RSTRIDE;
  const int split = blockIdx.x % nsplit, ch = blockIdx.x / nsplit;
  const int crop = ch / heads, head = ch % heads;
  const int width = heads * HD;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;
  const int P = L - 1;

  // stage Q, K, V of this head with cp.async (all copies in flight at once; padding rows are zeroed)
  for (int e = tid; e < LP * 8 * 3; e += AWARPS * 32) {  // 16-byte chunks: 8 per row per matrix
    const int c8 = e & 7, row = (e >> 3) % LP, mat = e / (8 * LP);
    const uint32_t dst = qt + (uint32_t)(mat * LP * RSTRIDE + row * RSTRIDE + c8 * 16);
    if (row < L) {
      const bf16* src = qkv + ((size_t)crop * L + row) * 3 * width + mat * width + head * HD + c8 * 8;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    } else {
      *reinterpret_cast<uint4*>(asmem + (size_t)mat * LP * RSTRIDE + row * RSTRIDE + c8 * 16) = make_uint4(0, 0, 0, 0);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const float scale = 0.125f;  // 64^-0.5
  const float* sim = simmap ? simmap + (size_t)crop * P * P : nullptr;
  int npass = 1;
  if (mode == CSEG_ATTN_SCLIP) npass = 2;
  if (mode == CSEG_ATTN_SEGEARTH) npass = 3;

  for (int rb = split * AWARPS + warp; rb * 16 < L; rb += AWARPS * nsplit) {
    const int r0 = rb * 16, row0 = r0 + g, row1 = r0 + g + 8;
    float O[8][4];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)
#pragma unroll
      for (int e = 0; e < 4; ++e) O[nb][e] = 0.f;
#pragma unroll 1
    for (int pass = 0; pass < npass; ++pass) {
      float S[NKB][4];
#pragma unroll
      for (int nb = 0; nb < NKB; ++nb)
#pragma unroll
        for (int e = 0; e < 4; ++e) S[nb][e] = 0.f;
      if (mode == CSEG_ATTN_STD || mode == CSEG_ATTN_VANILLA || mode == CSEG_ATTN_CAUSAL) {
        accum_scores<NKB>(S, qt, kt, r0, lane);
      } else if (mode == CSEG_ATTN_CLEARCLIP) {
        accum_scores<NKB>(S, qt, qt, r0, lane);
      } else if (mode == CSEG_ATTN_SFP || mode == CSEG_ATTN_EXPERIMENTAL) {
        accum_scores<NKB>(S, kt, kt, r0, lane);   // kk + qq into the same accumulator (:897-899)
        accum_scores<NKB>(S, qt, qt, r0, lane);
      } else {                                    // SCLIP / SegEarth: one self-similarity per pass
        const uint32_t t = pass == 0 ? qt : (pass == 1 ? kt : vt);
        accum_scores<NKB>(S, t, t, r0, lane);
      }
      const float sc = (mode == CSEG_ATTN_SFP) ? 0.5f * scale : scale;
#pragma unroll
      for (int nb = 0; nb < NKB; ++nb)
#pragma unroll
        for (int e = 0; e < 4; ++e) S[nb][e] *= sc;
      if (mode == CSEG_ATTN_EXPERIMENTAL) {
        frag_softmax<NKB>(S, L, tig);
        add_simmap<NKB>(S, sim, simw, P, L, row0, row1, tig);   // on the probabilities (:900-901)
        frag_softmax<NKB>(S, L, tig);                           // second softmax, unconditional (:902)
      } else {
        if (mode == CSEG_ATTN_CAUSAL) {           // additive -inf mask above the diagonal (build_attention_mask)
#pragma unroll
          for (int nb = 0; nb < NKB; ++nb)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int j = nb * 8 + 2 * tig + e;
              if (j > row0) S[nb][e] = -INFINITY;
              if (j > row1) S[nb][2 + e] = -INFINITY;
            }
        } else if (mode != CSEG_ATTN_STD) {
          add_simmap<NKB>(S, sim, simw, P, L, row0, row1, tig);
        }
        frag_softmax<NKB>(S, L, tig);
      }
      if (stats != nullptr) {
        float* st = stats + ((size_t)(crop * heads + head) * 2) * P;
#pragma unroll
        for (int nb = 0; nb < NKB; ++nb)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int j = nb * 8 + 2 * tig + e;
            if (j >= 1 && j < L) {
              if (row0 == 0) st[j - 1] = S[nb][e];
              if (row0 == j) st[P + j - 1] = S[nb][e];
              if (row1 == j) st[P + j - 1] = S[nb][2 + e];
            }
          }
      }
      // O += P . V : probability fragments of two adjacent key blocks form one A fragment
      const int q = lane >> 3, rr = lane & 7;
#pragma unroll
      for (int kb = 0; kb < NKB / 2; ++kb) {
        uint32_t a[4];
        a[0] = pack_bf16(S[2 * kb][0], S[2 * kb][1]);
        a[1] = pack_bf16(S[2 * kb][2], S[2 * kb][3]);
        a[2] = pack_bf16(S[2 * kb + 1][0], S[2 * kb + 1][1]);
        a[3] = pack_bf16(S[2 * kb + 1][2], S[2 * kb + 1][3]);
        // V^T fragments via ldmatrix.trans: matrices (keys 0-7, d blk 2j) (keys 8-15, d blk 2j) (keys 0-7, 2j+1) (8-15, 2j+1)
        const uint32_t base = vt + (uint32_t)((kb * 16 + (q & 1) * 8 + rr) * RSTRIDE + (q >> 1) * 16);
#pragma unroll
        for (int j2 = 0; j2 < 4; ++j2) {
          uint32_t b[4];
          ldsm_x4_t(base + j2 * 32, b);
          mma16816(O[2 * j2], a, b[0], b[1]);
          mma16816(O[2 * j2 + 1], a, b[2], b[3]);
        }
      }
    }
#pragma unroll
    for (int hm = 0; hm < 2; ++hm) {
      const int row = hm ? row1 : row0;
      if (row >= L) continue;
      bf16* o = out + ((size_t)crop * L + row) * width + head * HD + 2 * tig;
#pragma unroll
      for (int nb = 0; nb < 8; ++nb)
        *reinterpret_cast<__nv_bfloat162*>(o + nb * 8) = __floats2bfloat162_rn(O[nb][hm * 2], O[nb][hm * 2 + 1]);
    }
  }
}

template <int NKB>
int launch(const bf16* qkv, int n_crops, int L, int heads, int mode, const float* simmap, float simw, bf16* out,
           float* stats, cudaStream_t st) {
  const int smem = 3 * NKB * 8 * RSTRIDE;
  if (mode == CSEG_ATTN_STD) {
    CSEG_SET_SMEM((attention_mma_kernel<NKB, SWARPS, true>), smem);
    cseg_launch(attention_mma_kernel<NKB, SWARPS, true>, dim3(n_crops * heads), dim3(SWARPS * 32), smem, st, qkv, L, heads, mode,
                1, simmap, simw, out, stats);
  } else {
    CSEG_SET_SMEM((attention_mma_kernel<NKB, AWARPS, false>), smem);
    const int nsplit = (L + 16 * AWARPS - 1) / (16 * AWARPS);
    cseg_launch(attention_mma_kernel<NKB, AWARPS, false>, dim3(n_crops * heads * nsplit), dim3(AWARPS * 32), smem, st, qkv, L,
                heads, mode, nsplit, simmap, simw, out, stats);
  }
  CSEG_LAUNCH_CHECK("attention_mma");
  return 0;
}

}  // namespace

// returns 1 when the shape is not covered (caller falls back to the CUDA-core kernel)
int cseg_attention_mma(const bf16* qkv, int n_crops, int L, int heads, int head_dim, int mode, const float* simmap,
                       float simw, bf16* out, float* stats, cudaStream_t st) {
  if (head_dim != HD || mode == CSEG_ATTN_MASKCLIP || L > 272) return 1;
  if (L <= 80 && mode != CSEG_ATTN_STD) return launch<10>(qkv, n_crops, L, heads, mode, simmap, simw, out, stats, st);
  if (L <= 208) return launch<26>(qkv, n_crops, L, heads, mode, simmap, simw, out, stats, st);
  return launch<34>(qkv, n_crops, L, heads, mode, simmap, simw, out, stats, st);
}
