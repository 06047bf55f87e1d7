// Standard multi-head attention of the ViT blocks on tcgen05 (bf16, head_dim 64, L <= 272; described for L <= 208, the
// differences of the 272-key shape are listed at AtCfg):
//   out = softmax(q k^T / 8) v                                (nn.MultiheadAttention, open_clip/transformer.py:204,218-232)
//
// One work item = one (crop, head); a persistent CTA per SM walks the items.  L <= 208 keys fit one MMA in N, so there
// is no online softmax: per 128-query tile
//   S[128, 208]  = Q_tile . K^T       4 x tcgen05.mma (K = 16) into TMEM (two S buffers: the MMA warp runs a tile ahead)
//   P            = exp2((S - rowmax) * scale * log2 e)   by 16 warps: row quarter = warp % 4 (the TMEM lane window a warp
//                  may touch), key quarter = warp / 4; row max / row sum are exchanged through shared memory; P is written
//                  as bf16 straight into the K-major SWIZZLE_128B layout the second MMA reads
//   O[128, 64]   = P . V              13 x tcgen05.mma; V is consumed in its natural [key][dim] layout as an MN-major
//                  SWIZZLE_128B B operand (no transpose); 1 / rowsum is applied in the epilogue.
// Q / K / V tiles arrive by TMA from the [rows, 3*width] QKV matrix (the rows of a crop that lie beyond L belong to the
// next crop: those key columns are masked, those query rows are never stored).
// Warp roles: 0-15 softmax + epilogue, 16 TMA producer, 17 TMEM allocator + MMA issuer.
#include "common.cuh"
#include "tc_common.cuh"
#include <stdlib.h>

namespace {

constexpr int AT_HD = 64, AT_BM = 128, AT_LP = 208;           // keys padded to 13 K-steps of 16
constexpr int AT_SMW = 16;                                     // softmax / epilogue warps: row quarter = warp % 4, key quarter = warp / 4
constexpr int AT_THREADS = 32 * (AT_SMW + 2);
constexpr int AT_Q_BYTES = AT_BM * 128, AT_P_BLK = AT_BM * 128;
// (the shared-memory layout of the standard kernel depends on the padded key count: AtCfg<LP> below)
constexpr int AT_O_COL = 448, AT_TMEM_COLS = 512;
// key quarters: [0, 64), [64, 112), [112, 160), [160, 208)
__host__ __device__ constexpr int at_qbegin(int k) { return k == 0 ? 0 : (k == 1 ? 64 : (k == 2 ? 112 : (k == 3 ? 160 : AT_LP))); }

// Shape of the standard kernel by padded key count LP.  LP = 208 (ViT-B/16 at 224: L = 197): two S accumulators, K and V
// double buffered.  LP = 272 (ViT-L/14 at 224: L = 257): 272 + 64 TMEM columns leave room for ONE S accumulator (the next
// S = Q.K^T is issued as soon as the softmax warps have read the current one), S is issued as two MMAs (N <= 256), K / V
// arrive as two TMA boxes (box rows <= 256), K stays double buffered and V single (227 KB of shared memory): V of the next
// item is only needed one softmax after its K.
template <int LP>
struct AtCfg {
  static_assert(LP % 16 == 0 && LP <= 272, "keys padded to k-steps of 16");
  static constexpr int NSB = LP <= 224 ? 2 : 1;                 // S accumulators in TMEM
  static constexpr int S_COLS = 224;                            // column stride between them
  static constexpr int NVB = LP <= 208 ? 2 : 1;                 // V buffers
  static constexpr int KV_BYTES = LP * 128;
  static constexpr int KV_BOX = LP <= 256 ? LP : LP / 2;        // TMA box rows
  static constexpr int KV_LOADS = LP / KV_BOX;
  static constexpr int N0 = LP <= 256 ? LP : 144, N1 = LP - N0; // S = [N0 | N1] key columns per MMA
  static constexpr int P_BYTES = ((LP + 63) / 64) * AT_P_BLK;
  static constexpr int Q_OFF = 0, K_OFF = 2 * AT_Q_BYTES, V_OFF = K_OFF + 2 * KV_BYTES, P_OFF = V_OFF + NVB * KV_BYTES;
  static constexpr int RED_OFF = P_OFF + P_BYTES, RED_BYTES = (4 * 128 + 2 * 4 * 128 + LP) * 4;
  static constexpr int BAR_OFF = RED_OFF + RED_BYTES, NBARS = 19;
  static constexpr int SMEM = BAR_OFF + NBARS * 8 + 16 + 1024;
  static_assert(KV_BOX % 8 == 0 && N0 % 16 == 0 && N1 % 16 == 0 && (N0 * 128) % 1024 == 0, "swizzle atoms");
  static_assert(SMEM <= 227 * 1024, "shared memory");
  // key quarters of the softmax warps (multiples of 16)
  __host__ __device__ static constexpr int qbegin(int k) {
    return LP <= 208 ? at_qbegin(k) : (k == 0 ? 0 : (k == 1 ? 80 : (k == 2 ? 144 : (k == 3 ? 208 : LP))));
  }
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float at_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t at_pack(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
// MN-major SWIZZLE_128B descriptor (V as the B operand of P.V: 8 keys x 64 dims per 1024 B atom)
__device__ __forceinline__ uint64_t at_mn_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(1024 >> 4) << 16;       // LBO: unused (N = 64 is one atom wide)
  d |= (uint64_t)(1024 >> 4) << 32;       // SBO: stride between 8-key groups
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16, D = f32, A = B = bf16, A K-major, B MN-major (bit 16)
__host__ __device__ constexpr uint32_t at_idesc_pv() {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(AT_HD >> 3) << 17) | ((uint32_t)(AT_BM >> 4) << 24);
}

// STATS: additionally emits P[0, 1+i] and P[1+i, 1+i] per (crop, head) -- the entries of the need_weights=True matrix that
// detect_outliers_by_attention reads (outlier_suppression.py:46-53): the diagonal numerator stays in a register of the
// thread that owns row i, the CLS row's numerators are parked in shared memory; both are normalised in the epilogue.
template <bool STATS, int LP>
__global__ void __launch_bounds__(AT_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, int L, int heads,
                    int n_items, int mt, bf16* __restrict__ out, float scale_log2e, int diag, float* __restrict__ stats) {
  using C = AtCfg<LP>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic on the __shared__ array: accesses compile to LDS / STS, not generic LD / ST
  const uint32_t sbase = smem_u32(smem);
  float* smax = reinterpret_cast<float*>(smem + C::RED_OFF);           // [4][128]
  float* ssum = smax + 4 * 128;                                        // [2 tiles][4][128]
  float* scls = ssum + 2 * 4 * 128;                                    // [LP] numerators of the CLS row (STATS)
  uint64_t* bars = (uint64_t*)(smem + C::BAR_OFF);
  uint32_t* tmem_slot = (uint32_t*)(bars + C::NBARS);
  const uint32_t b0 = smem_u32(bars);
  const uint32_t q_full = b0, q_empty = b0 + 16, k_full = b0 + 32, k_empty = b0 + 48, v_full = b0 + 64, v_empty = b0 + 80;
  const uint32_t s_full = b0 + 96, s_empty = b0 + 112, p_full = b0 + 128, o_full = b0 + 136, o_empty = b0 + 144;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int width = heads * AT_HD;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmKV) : "memory");
    for (int i = 0; i < 2; ++i) {
      mbar_init(q_full + i * 8, 1);
      mbar_init(q_empty + i * 8, 1);
      mbar_init(k_full + i * 8, 1);
      mbar_init(k_empty + i * 8, 1);
      mbar_init(v_full + i * 8, 1);
      mbar_init(v_empty + i * 8, 1);
      mbar_init(s_full + i * 8, 1);
      mbar_init(s_empty + i * 8, AT_SMW);
    }
    mbar_init(p_full, AT_SMW);
    mbar_init(o_full, 1);
    mbar_init(o_empty, AT_SMW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == AT_SMW + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)AT_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_grid_sync();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == AT_SMW) {
    // ---------------- TMA producer: K, Q tile 0, V, the other Q tiles (V's buffer may free up a softmax later than K's) ------
    if (lane == 0) {
      uint32_t qi = 0, ki = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ki) {
        const int crop = item / heads, head = item - crop * heads;
        const uint32_t kb = ki & 1, kph = (ki >> 1) & 1, vb = ki % C::NVB, vph = (ki / C::NVB) & 1;
        mbar_wait(k_empty + kb * 8, kph ^ 1);
        mbar_expect_tx(k_full + kb * 8, C::KV_BYTES);
#pragma unroll
        for (int l = 0; l < C::KV_LOADS; ++l)
          tma_load_2d(sbase + C::K_OFF + kb * C::KV_BYTES + l * C::KV_BOX * 128, &tmKV, k_full + kb * 8, width + head * AT_HD,
                      crop * L + l * C::KV_BOX);
        for (int t = 0; t < mt; ++t, ++qi) {
          const uint32_t qb = qi & 1, qph = (qi >> 1) & 1;
          mbar_wait(q_empty + qb * 8, qph ^ 1);
          mbar_expect_tx(q_full + qb * 8, AT_Q_BYTES);
          tma_load_2d(sbase + C::Q_OFF + qb * AT_Q_BYTES, &tmQ, q_full + qb * 8, head * AT_HD, crop * L + t * AT_BM);
          if (t == 0) {
            mbar_wait(v_empty + vb * 8, vph ^ 1);
            mbar_expect_tx(v_full + vb * 8, C::KV_BYTES);
#pragma unroll
            for (int l = 0; l < C::KV_LOADS; ++l)
              tma_load_2d(sbase + C::V_OFF + vb * C::KV_BYTES + l * C::KV_BOX * 128, &tmKV, v_full + vb * 8,
                          2 * width + head * AT_HD, crop * L + l * C::KV_BOX);
          }
        }
      }
    }
  } else if (warp == AT_SMW + 1) {
    // ---------------- MMA issuer: S of tile i is issued before P.V of tile i-1 ----------------
    constexpr uint32_t idesc_s0 = make_idesc(AT_BM, C::N0), idesc_s1 = make_idesc(AT_BM, C::N1 > 0 ? C::N1 : 16),
                       idesc_pv = at_idesc_pv();
    uint32_t qi = 0, ki = 0;
    bool have_prev = false;
    uint32_t prev_pi = 0, prev_vb = 0, prev_vph = 0;
    bool prev_first = false, prev_last = false;
    auto issue_pv = [&]() {
      if (prev_first) mbar_wait(v_full + prev_vb * 8, prev_vph);
      mbar_wait(p_full, prev_pi & 1);
      mbar_wait(o_empty, (prev_pi & 1) ^ 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < LP / 16; ++j) {
          const uint64_t adesc = make_sdesc(sbase + C::P_OFF + (j >> 2) * AT_P_BLK + (j & 3) * 32);
          const uint64_t bdesc = at_mn_desc(sbase + C::V_OFF + prev_vb * C::KV_BYTES + j * 2048);
          if (!(diag & 2)) umma_f16(tmem_base + AT_O_COL, adesc, bdesc, idesc_pv, j > 0 ? 1u : 0u);
        }
        umma_commit(o_full);
        if (prev_last) umma_commit(v_empty + prev_vb * 8);
      }
      __syncwarp();
    };
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ki) {
      const uint32_t kb = ki & 1, kph = (ki >> 1) & 1, vb = ki % C::NVB, vph = (ki / C::NVB) & 1;
      for (int t = 0; t < mt; ++t, ++qi) {
        const uint32_t qb = qi & 1, qph = (qi >> 1) & 1;
        const uint32_t sb = C::NSB == 2 ? qb : 0u, sph = C::NSB == 2 ? qph : (qi & 1);
        if (t == 0) mbar_wait(k_full + kb * 8, kph);
        mbar_wait(q_full + qb * 8, qph);
        mbar_wait(s_empty + sb * 8, sph ^ 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < AT_HD / 16; ++ks) {
            const uint64_t adesc = make_sdesc(sbase + C::Q_OFF + qb * AT_Q_BYTES + ks * 32);
            const uint64_t bdesc = make_sdesc(sbase + C::K_OFF + kb * C::KV_BYTES + ks * 32);
            if (!(diag & 4)) {
              umma_f16(tmem_base + sb * C::S_COLS, adesc, bdesc, idesc_s0, ks > 0 ? 1u : 0u);
              if (C::N1 > 0)
                umma_f16(tmem_base + sb * C::S_COLS + C::N0, adesc,
                         make_sdesc(sbase + C::K_OFF + kb * C::KV_BYTES + C::N0 * 128 + ks * 32), idesc_s1, ks > 0 ? 1u : 0u);
            }
          }
          umma_commit(s_full + sb * 8);
          umma_commit(q_empty + qb * 8);
          if (t == mt - 1) umma_commit(k_empty + kb * 8);
        }
        __syncwarp();
        if (have_prev) issue_pv();
        have_prev = true;
        prev_pi = qi;
        prev_vb = vb;
        prev_vph = vph;
        prev_first = (t == 0);
        prev_last = (t == mt - 1);
      }
    }
    if (have_prev) issue_pv();
  } else {
    // ---------------- softmax + epilogue ----------------
    const int q4 = warp & 3, hh = warp >> 2, row = q4 * 32 + lane;         // hh: key quarter
    const int c_begin = C::qbegin(hh), c_end = C::qbegin(hh + 1);
    const uint32_t tlane = tmem_base + ((uint32_t)(q4 * 32) << 16);
    uint8_t* prow = smem + C::P_OFF + row * 128;
    uint32_t qi = 0;
    bool have_prev = false;
    uint32_t prev_pi = 0;
    int prev_row0 = 0, prev_head = 0, prev_valid = 0, prev_t = 0, prev_item = 0;
    float ediag = 0.f;                                                        // STATS: numerator of P[tok, tok]
    auto epilogue = [&]() {
      mbar_wait(o_full, prev_pi & 1);
      tc_fence_after();
      uint32_t r[16];
      tmem_ld16(tlane + AT_O_COL + hh * 16, r);
      const float* ss = ssum + (prev_pi & 1) * 512;
      const float inv = 1.0f / ((ss[row] + ss[128 + row]) + (ss[256 + row] + ss[384 + row]));
      if (STATS) {
        const int P = L - 1, ptok = prev_t * AT_BM + row;
        float* st = stats + (size_t)prev_item * 2 * P;
        if (ptok >= 1 && ptok < L && ptok >= c_begin && ptok < c_end) st[P + ptok - 1] = ediag * inv;
        if (prev_t == 0 && q4 == 0) {                                        // CLS row: this warp's key quarter, lane-parallel
          const float inv0 = __shfl_sync(0xffffffffu, inv, 0);
          for (int c = c_begin + lane; c < c_end; c += 32)
            if (c >= 1 && c < L) st[c - 1] = scls[c] * inv0;
        }
      }
      if (row < prev_valid && !(diag & 8)) {
        uint4* o = reinterpret_cast<uint4*>(out + (size_t)(prev_row0 + row) * width + prev_head * AT_HD + hh * 16);
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          uint4 w;
          w.x = at_pack(__uint_as_float(r[8 * v]) * inv, __uint_as_float(r[8 * v + 1]) * inv);
          w.y = at_pack(__uint_as_float(r[8 * v + 2]) * inv, __uint_as_float(r[8 * v + 3]) * inv);
          w.z = at_pack(__uint_as_float(r[8 * v + 4]) * inv, __uint_as_float(r[8 * v + 5]) * inv);
          w.w = at_pack(__uint_as_float(r[8 * v + 6]) * inv, __uint_as_float(r[8 * v + 7]) * inv);
          o[v] = w;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
    };
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int crop = item / heads, head = item - crop * heads;
      for (int t = 0; t < mt; ++t, ++qi) {
        const uint32_t sb = C::NSB == 2 ? (qi & 1) : 0u, sph = C::NSB == 2 ? ((qi >> 1) & 1) : (qi & 1);
        mbar_wait(s_full + sb * 8, sph);
        tc_fence_after();
        const uint32_t ts = tlane + sb * C::S_COLS;
        // pass 1: row maximum over this warp's key half, then across the two halves.  The TMEM load of the next 16
        // columns is in flight while the current 16 are reduced (the softmax warps were latency bound: ncu shows
        // long-scoreboard stalls on a tcgen05.ld + wait per chunk with only two warps per scheduler)
        float m = -INFINITY;
        {
          uint32_t ra[16], rb[16];
          auto red = [&](const uint32_t (&r)[16], int c) {
            if (c + 16 <= L) {               // warp-uniform: only the chunk that holds key L needs per-element masks
#pragma unroll
              for (int e = 0; e < 16; ++e) m = fmaxf(m, __uint_as_float(r[e]));
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e)
                if (c + e < L) m = fmaxf(m, __uint_as_float(r[e]));
            }
          };
          tmem_ld16_nowait(ts + c_begin, ra);
#pragma unroll 1
          for (int c = c_begin; c < ((diag & 16) ? c_begin + 1 : c_end); c += 32) {
            tmem_wait_ld();
            if (c + 16 < c_end) tmem_ld16_nowait(ts + c + 16, rb);
            red(ra, c);
            if (c + 16 < c_end) {
              tmem_wait_ld();
              if (c + 32 < c_end) tmem_ld16_nowait(ts + c + 32, ra);
              red(rb, c + 16);
            }
          }
        }
        smax[hh * 128 + row] = m;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * AT_SMW) : "memory");
        m = fmaxf(fmaxf(smax[row], smax[128 + row]), fmaxf(smax[256 + row], smax[384 + row]));
        // epilogue of the previous tile: its P.V has completed, so the P buffer is free for this tile
        if (have_prev) epilogue();
        // pass 2: exponentials -> bf16 P in the K-major SWIZZLE_128B layout, row sum in fp32
        const float msc = m * scale_log2e;
        float sum = 0.f;
        {
          uint32_t ra[16], rb[16];
          auto expo = [&](const uint32_t (&r)[16], int c) {
            float ev[16];
            if (c + 16 <= L) {
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                ev[e] = at_ex2(fmaf(__uint_as_float(r[e]), scale_log2e, -msc));
                sum += ev[e];
              }
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                ev[e] = (c + e < L) ? at_ex2(fmaf(__uint_as_float(r[e]), scale_log2e, -msc)) : 0.f;
                sum += ev[e];
              }
            }
            if (STATS) {
              const int tok = t * AT_BM + row;
#pragma unroll
              for (int e = 0; e < 16; ++e)
                if (c + e == tok) ediag = ev[e];
              if (tok == 0) {
#pragma unroll
                for (int e = 0; e < 16; ++e) scls[c + e] = ev[e];
              }
            }
            uint8_t* pb = prow + (c >> 6) * AT_P_BLK;
            const int j = (c & 63) >> 3;                                   // 16-byte chunk within the 128-byte row (even)
            uint4 w0, w1;
            w0.x = at_pack(ev[0], ev[1]);   w0.y = at_pack(ev[2], ev[3]);   w0.z = at_pack(ev[4], ev[5]);   w0.w = at_pack(ev[6], ev[7]);
            w1.x = at_pack(ev[8], ev[9]);   w1.y = at_pack(ev[10], ev[11]); w1.z = at_pack(ev[12], ev[13]); w1.w = at_pack(ev[14], ev[15]);
            *reinterpret_cast<uint4*>(pb + ((j ^ (row & 7)) << 4)) = w0;
            *reinterpret_cast<uint4*>(pb + (((j + 1) ^ (row & 7)) << 4)) = w1;
          };
          tmem_ld16_nowait(ts + c_begin, ra);
#pragma unroll 1
          for (int c = c_begin; c < ((diag & 1) ? c_begin + 1 : c_end); c += 32) {
            tmem_wait_ld();
            if (c + 16 < c_end) tmem_ld16_nowait(ts + c + 16, rb);
            expo(ra, c);
            if (c + 16 < c_end) {
              tmem_wait_ld();
              if (c + 32 < c_end) tmem_ld16_nowait(ts + c + 32, ra);
              expo(rb, c + 16);
            }
          }
        }
        ssum[(qi & 1) * 512 + hh * 128 + row] = sum;
        tc_fence_before();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(p_full);
          mbar_arrive(s_empty + sb * 8);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * AT_SMW) : "memory");
        have_prev = true;
        prev_pi = qi;
        prev_row0 = crop * L + t * AT_BM;
        prev_head = head;
        prev_valid = min(AT_BM, L - t * AT_BM);
        prev_t = t;
        prev_item = item;
      }
    }
    if (have_prev) epilogue();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == AT_SMW + 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)AT_TMEM_COLS) : "memory");
  }
}

// ----------------------------------------------------------------------------------------------------------------
// Final block, model_type 'Experimental' (open_clip/transformer.py:897-903) on the same pipeline:
//   P = softmax(softmax((k k^T + q q^T) / 8) + w * M_pad),  out = P v
// S: 8 MMAs per 128-query tile (K_tile . K^T, then Q_tile . Q^T into the same accumulator); the A operand of a tile is a
// row window of the 208-row K / Q buffer that also serves as the B operand.  Three passes over S in TMEM: (1) row max,
// (2) first-softmax sum + row max of w * M (the similarity map, zero CLS row / column), (3) second softmax numerators
// exp(p1 + w M - (max(w M, 0) + 1)) -- the offset is within 1 of the true row maximum, so nothing over- or underflows --
// written as bf16 P.  Row statistics of the four key quarters are exchanged through shared memory.
// smem: Q-all x2, K-all x2, V x1 (V of the next item is only needed one softmax later), P, reductions.
// ----------------------------------------------------------------------------------------------------------------
// LP = 272 (ViT-L/14 crops): as AtCfg -- one S accumulator, S as two MMAs, two TMA boxes per matrix -- and ONE Q-all / K-all
// buffer pair (the next item's Q / K are fetched while the last tile's softmax and P.V run).
template <int LP>
struct AxCfg {
  using A = AtCfg<LP>;
  static constexpr int NQKB = LP <= 208 ? 2 : 1;
  static constexpr int QK_OFF = 0, V_OFF = 2 * NQKB * A::KV_BYTES, P_OFF = V_OFF + A::KV_BYTES, RED_OFF = P_OFF + A::P_BYTES;
  static constexpr int RED_BYTES = (3 * 4 * 128 + 2 * 4 * 128) * 4;   // smax[4][128], ssum1[4][128], smaxm[4][128], ssum[2][4][128]
  static constexpr int BAR_OFF = RED_OFF + RED_BYTES, NBARS = 13;
  static constexpr int SMEM = BAR_OFF + NBARS * 8 + 16 + 1024;
  static constexpr int SIM_COLS = LP, SIM_FLOATS = (LP + 31) / 32 * LP * 32;   // layout 1 of cseg_simmap_tc
  static_assert(SMEM <= 227 * 1024, "shared memory");
};

template <int LP>
__global__ void __launch_bounds__(AT_THREADS, 1)
attention_tc_exp_kernel(const __grid_constant__ CUtensorMap tmKV, int L, int heads, int n_items, bf16* __restrict__ out,
                        float scale_log2e, const float* __restrict__ simmap, float simw) {
  using C = AxCfg<LP>;
  using A = AtCfg<LP>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic on the __shared__ array: accesses compile to LDS / STS, not generic LD / ST
  const uint32_t sbase = smem_u32(smem);
  float* smax = reinterpret_cast<float*>(smem + C::RED_OFF);           // [4][128]
  float* ssum1 = smax + 4 * 128;                                       // [4][128]
  float* smaxm = ssum1 + 4 * 128;                                      // [4][128]
  float* ssum = smaxm + 4 * 128;                                       // [2][4][128]
  uint64_t* bars = (uint64_t*)(smem + C::BAR_OFF);
  uint32_t* tmem_slot = (uint32_t*)(bars + C::NBARS);
  const uint32_t b0 = smem_u32(bars);
  const uint32_t qk_full = b0, qk_empty = b0 + 16, v_full = b0 + 32, v_empty = b0 + 40, s_full = b0 + 48, s_empty = b0 + 64;
  const uint32_t p_full = b0 + 80, o_full = b0 + 88, o_empty = b0 + 96;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int width = heads * AT_HD, mt = (L + AT_BM - 1) / AT_BM;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmKV) : "memory");
    for (int i = 0; i < 2; ++i) {
      mbar_init(qk_full + i * 8, 1);
      mbar_init(qk_empty + i * 8, 1);
      mbar_init(s_full + i * 8, 1);
      mbar_init(s_empty + i * 8, AT_SMW);
    }
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_init(p_full, AT_SMW);
    mbar_init(o_full, 1);
    mbar_init(o_empty, AT_SMW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == AT_SMW + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)AT_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_grid_sync();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == AT_SMW) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      uint32_t ki = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ki) {
        const int crop = item / heads, head = item - crop * heads;
        const uint32_t kb = ki % C::NQKB, kph = (ki / C::NQKB) & 1;
        mbar_wait(qk_empty + kb * 8, kph ^ 1);
        mbar_expect_tx(qk_full + kb * 8, 2 * A::KV_BYTES);
#pragma unroll
        for (int l = 0; l < A::KV_LOADS; ++l) {
          tma_load_2d(sbase + C::QK_OFF + (2 * kb) * A::KV_BYTES + l * A::KV_BOX * 128, &tmKV, qk_full + kb * 8, head * AT_HD,
                      crop * L + l * A::KV_BOX);
          tma_load_2d(sbase + C::QK_OFF + (2 * kb + 1) * A::KV_BYTES + l * A::KV_BOX * 128, &tmKV, qk_full + kb * 8,
                      width + head * AT_HD, crop * L + l * A::KV_BOX);
        }
        mbar_wait(v_empty, (ki & 1) ^ 1);
        mbar_expect_tx(v_full, A::KV_BYTES);
#pragma unroll
        for (int l = 0; l < A::KV_LOADS; ++l)
          tma_load_2d(sbase + C::V_OFF + l * A::KV_BOX * 128, &tmKV, v_full, 2 * width + head * AT_HD, crop * L + l * A::KV_BOX);
      }
    }
  } else if (warp == AT_SMW + 1) {
    // ---------------- MMA issuer: S of tile i is issued before P.V of tile i-1 ----------------
    constexpr uint32_t idesc_s0 = make_idesc(AT_BM, A::N0), idesc_s1 = make_idesc(AT_BM, A::N1 > 0 ? A::N1 : 16),
                       idesc_pv = at_idesc_pv();
    uint32_t qi = 0, ki = 0;
    bool have_prev = false;
    uint32_t prev_pi = 0, prev_ki = 0;
    bool prev_first = false, prev_last = false;
    auto issue_pv = [&]() {
      mbar_wait(p_full, prev_pi & 1);
      mbar_wait(o_empty, (prev_pi & 1) ^ 1);
      if (prev_first) mbar_wait(v_full, prev_ki & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < LP / 16; ++j) {
          const uint64_t adesc = make_sdesc(sbase + C::P_OFF + (j >> 2) * AT_P_BLK + (j & 3) * 32);
          const uint64_t bdesc = at_mn_desc(sbase + C::V_OFF + j * 2048);
          umma_f16(tmem_base + AT_O_COL, adesc, bdesc, idesc_pv, j > 0 ? 1u : 0u);
        }
        umma_commit(o_full);
        if (prev_last) umma_commit(v_empty);
      }
      __syncwarp();
    };
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ki) {
      const uint32_t kb = ki % C::NQKB, kph = (ki / C::NQKB) & 1;
      for (int t = 0; t < mt; ++t, ++qi) {
        const uint32_t qb = A::NSB == 2 ? (qi & 1) : 0u, qph = A::NSB == 2 ? ((qi >> 1) & 1) : (qi & 1);   // S buffer / phase
        if (t == 0) mbar_wait(qk_full + kb * 8, kph);
        mbar_wait(s_empty + qb * 8, qph ^ 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int mat = 1; mat >= 0; --mat) {                       // k k^T first, then q q^T (transformer.py:897-899)
            const uint32_t mbase = sbase + C::QK_OFF + (2 * kb + mat) * A::KV_BYTES;
#pragma unroll
            for (int ks = 0; ks < AT_HD / 16; ++ks) {
              const uint64_t adesc = make_sdesc(mbase + t * AT_Q_BYTES + ks * 32);
              umma_f16(tmem_base + qb * A::S_COLS, adesc, make_sdesc(mbase + ks * 32), idesc_s0, (mat == 0 || ks > 0) ? 1u : 0u);
              if (A::N1 > 0)
                umma_f16(tmem_base + qb * A::S_COLS + A::N0, adesc, make_sdesc(mbase + A::N0 * 128 + ks * 32), idesc_s1,
                         (mat == 0 || ks > 0) ? 1u : 0u);
            }
          }
          umma_commit(s_full + qb * 8);
          if (t == mt - 1) umma_commit(qk_empty + kb * 8);
        }
        __syncwarp();
        if (have_prev) issue_pv();
        have_prev = true;
        prev_pi = qi;
        prev_ki = ki;
        prev_first = (t == 0);
        prev_last = (t == mt - 1);
      }
    }
    if (have_prev) issue_pv();
  } else {
    // ---------------- softmax + epilogue ----------------
    const int q4 = warp & 3, hh = warp >> 2, row = q4 * 32 + lane;         // hh: key quarter
    const int c_begin = A::qbegin(hh), c_end = A::qbegin(hh + 1);
    const uint32_t tlane = tmem_base + ((uint32_t)(q4 * 32) << 16);
    uint8_t* prow = smem + C::P_OFF + row * 128;
    const float LOG2E = 1.4426950408889634f;
    uint32_t qi = 0;
    bool have_prev = false;
    uint32_t prev_pi = 0;
    int prev_row0 = 0, prev_head = 0, prev_valid = 0;
    auto epilogue = [&]() {
      mbar_wait(o_full, prev_pi & 1);
      tc_fence_after();
      uint32_t r[16];
      tmem_ld16(tlane + AT_O_COL + hh * 16, r);
      const float* ss = ssum + (prev_pi & 1) * 512;
      const float inv = 1.0f / ((ss[row] + ss[128 + row]) + (ss[256 + row] + ss[384 + row]));
      if (row < prev_valid) {
        uint4* o = reinterpret_cast<uint4*>(out + (size_t)(prev_row0 + row) * width + prev_head * AT_HD + hh * 16);
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          uint4 w;
          w.x = at_pack(__uint_as_float(r[8 * v]) * inv, __uint_as_float(r[8 * v + 1]) * inv);
          w.y = at_pack(__uint_as_float(r[8 * v + 2]) * inv, __uint_as_float(r[8 * v + 3]) * inv);
          w.z = at_pack(__uint_as_float(r[8 * v + 4]) * inv, __uint_as_float(r[8 * v + 5]) * inv);
          w.w = at_pack(__uint_as_float(r[8 * v + 6]) * inv, __uint_as_float(r[8 * v + 7]) * inv);
          o[v] = w;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
    };
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int crop = item / heads, head = item - crop * heads;
      for (int t = 0; t < mt; ++t, ++qi) {
        const uint32_t sb = A::NSB == 2 ? (qi & 1) : 0u, sph = A::NSB == 2 ? ((qi >> 1) & 1) : (qi & 1);
        const int tok = t * AT_BM + row;                                      // token index of this thread's query row
        // similarity row of this query: M_pad[tok][j] = M[tok-1][j-1] for tok, j >= 1, else 0 (similarity_enhancement.py:104-107)
        // (layout 1 of cseg_simmap_tc: [crop][token / 32][key][token % 32]: a warp's 32 rows of one key are one 128-byte line;
        // the CLS row / column hold zeros, rows beyond L are never used: they read row 0)
        const bool has_sim = simmap != nullptr;                                  // uniform
        const int tokc = tok < L ? tok : 0;
        const float* mrow = has_sim ? simmap + (size_t)crop * C::SIM_FLOATS + (size_t)(tokc >> 5) * (C::SIM_COLS * 32) + (tokc & 31)
                                    : nullptr;
        mbar_wait(s_full + sb * 8, sph);
        tc_fence_after();
        const uint32_t ts = tlane + sb * A::S_COLS;
        uint32_t r[16];
        // pass 1: row maximum of S
        float m = -INFINITY;
#pragma unroll 1
        for (int c = c_begin; c < c_end; c += 16) {
          tmem_ld16(ts + c, r);
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (c + e < L) m = fmaxf(m, __uint_as_float(r[e]));
        }
        smax[hh * 128 + row] = m;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * AT_SMW) : "memory");
        m = fmaxf(fmaxf(smax[row], smax[128 + row]), fmaxf(smax[256 + row], smax[384 + row]));
        if (have_prev) epilogue();       // P.V of the previous tile has completed: the P buffer is free
        // pass 2: sum of the first softmax, row maximum of w * M
        const float msc = m * scale_log2e;
        float sum1 = 0.f, mm = 0.f;       // M_pad holds zeros (CLS column): 0 takes part in the maximum
#pragma unroll 1
        for (int c = c_begin; c < c_end; c += 16) {
          float mv[16];                   // all 16 loads in flight before the TMEM wait (one coalesced line per key)
          if (has_sim) {
#pragma unroll
            for (int e = 0; e < 16; ++e) mv[e] = __ldg(mrow + (c + e) * 32);
          }
          tmem_ld16(ts + c, r);
          if (c + 16 <= L) {
#pragma unroll
            for (int e = 0; e < 16; ++e) sum1 += at_ex2(fmaf(__uint_as_float(r[e]), scale_log2e, -msc));
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if (c + e < L) sum1 += at_ex2(fmaf(__uint_as_float(r[e]), scale_log2e, -msc));
          }
          if (has_sim) {
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if (c + e < L) mm = fmaxf(mm, simw * mv[e]);
          }
        }
        ssum1[hh * 128 + row] = sum1;
        smaxm[hh * 128 + row] = mm;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * AT_SMW) : "memory");
        const float inv1 = 1.0f / ((ssum1[row] + ssum1[128 + row]) + (ssum1[256 + row] + ssum1[384 + row]));
        const float off = (fmaxf(fmaxf(smaxm[row], smaxm[128 + row]), fmaxf(smaxm[256 + row], smaxm[384 + row])) + 1.0f) * LOG2E;
        // pass 3: P = exp(p1 + w M - offset) as bf16 in the K-major SWIZZLE_128B layout, row sum in fp32
        float sum = 0.f;
#pragma unroll 1
        for (int c = c_begin; c < c_end; c += 16) {
          float ev[16];
          if (has_sim) {
#pragma unroll
            for (int e = 0; e < 16; ++e) ev[e] = simw * __ldg(mrow + (c + e) * 32);
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) ev[e] = 0.f;
          }
          tmem_ld16(ts + c, r);
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float tt = fmaf(at_ex2(fmaf(__uint_as_float(r[e]), scale_log2e, -msc)), inv1, ev[e]);
            const float v = at_ex2(fmaf(tt, LOG2E, -off));
            ev[e] = (c + e < L) ? v : 0.f;
            sum += ev[e];
          }
          uint8_t* pb = prow + (c >> 6) * AT_P_BLK;
          const int j = (c & 63) >> 3;                                   // 16-byte chunk within the 128-byte row (even)
          uint4 w0, w1;
          w0.x = at_pack(ev[0], ev[1]);   w0.y = at_pack(ev[2], ev[3]);   w0.z = at_pack(ev[4], ev[5]);   w0.w = at_pack(ev[6], ev[7]);
          w1.x = at_pack(ev[8], ev[9]);   w1.y = at_pack(ev[10], ev[11]); w1.z = at_pack(ev[12], ev[13]); w1.w = at_pack(ev[14], ev[15]);
          *reinterpret_cast<uint4*>(pb + ((j ^ (row & 7)) << 4)) = w0;
          *reinterpret_cast<uint4*>(pb + (((j + 1) ^ (row & 7)) << 4)) = w1;
        }
        ssum[(qi & 1) * 512 + hh * 128 + row] = sum;
        tc_fence_before();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(p_full);
          mbar_arrive(s_empty + sb * 8);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * AT_SMW) : "memory");
        have_prev = true;
        prev_pi = qi;
        prev_row0 = crop * L + t * AT_BM;
        prev_head = head;
        prev_valid = min(AT_BM, L - t * AT_BM);
      }
    }
    if (have_prev) epilogue();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == AT_SMW + 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)AT_TMEM_COLS) : "memory");
  }
}

typedef CUresult (*AtEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
AtEncodeFn at_encode() {
  static AtEncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (AtEncodeFn)p;
  }
  return fn;
}
// 2D bf16 map over the [rows, cols] QKV matrix: box {64 columns, box_rows}, SWIZZLE_128B, zero fill beyond the last row
int at_make_map(CUtensorMap* m, const void* base, long long rows, int cols, int box_rows) {
  AtEncodeFn enc = at_encode();
  if (!enc) CSEG_FAIL(CSEG_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) CSEG_FAIL(CSEG_ECUDA, "attention: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

bool at_enabled() {   // CSEG_ATTN_TC=0 selects the mma.sync kernel (A/B measurements)
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("CSEG_ATTN_TC");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

}  // namespace

template <bool STATS, int LP>
int at_launch(const bf16* qkv, int n_crops, int L, int heads, bf16* out, float* stats, cudaStream_t st) {
  using C = AtCfg<LP>;
  const int width = heads * AT_HD, items = n_crops * heads;
  CUtensorMap tq, tkv;
  if (int rc = at_make_map(&tq, qkv, (long long)n_crops * L, 3 * width, AT_BM)) return rc;
  if (int rc = at_make_map(&tkv, qkv, (long long)n_crops * L, 3 * width, C::KV_BOX)) return rc;
  CSEG_SET_SMEM((attention_tc_kernel<STATS, LP>), C::SMEM);
  // L = 257 leaves ONE query row for a third tile.  A separate kernel for such leftover rows (a CTA per row, K / V from L2)
  // was measured at 100 us per launch against 52 us for the extra tile (162 crops x 16 heads: 232 vs 185 us): not kept.
  const int mt = (L + AT_BM - 1) / AT_BM;
  const int grid = std::min(items, sm_count());
  const float scale_log2e = 0.125f * 1.4426950408889634f;     // head_dim^-0.5 * log2(e)
  static int diag = -1;                     // CSEG_ATTN_DIAG: knock-out bits for timing experiments (results invalid)
  if (diag < 0) {
    const char* e = getenv("CSEG_ATTN_DIAG");
    diag = e ? atoi(e) : 0;
  }
  cseg_launch(attention_tc_kernel<STATS, LP>, dim3(grid), dim3(AT_THREADS), C::SMEM, st, tq, tkv, L, heads, items, mt, out,
              scale_log2e, diag, stats);
  CSEG_LAUNCH_CHECK("attention_tc");
  return 0;
}

// returns 1 when the case is not covered (the caller falls back to the mma.sync kernel)
int cseg_attention_tc(const bf16* qkv, int n_crops, int L, int heads, int head_dim, int mode, const float* simmap, float simw,
                      bf16* out, float* stats, cudaStream_t st) {
  if (!at_enabled() || head_dim != AT_HD) return 1;
  if (mode != CSEG_ATTN_STD || simmap != nullptr) return 1;
  if (L < 17 || L > 272 || ((uintptr_t)qkv & 15) != 0 || ((uintptr_t)out & 15) != 0) return 1;
  if (L <= AT_LP)
    return stats != nullptr ? at_launch<true, AT_LP>(qkv, n_crops, L, heads, out, stats, st)
                            : at_launch<false, AT_LP>(qkv, n_crops, L, heads, out, stats, st);
  return stats != nullptr ? at_launch<true, 272>(qkv, n_crops, L, heads, out, stats, st)
                          : at_launch<false, 272>(qkv, n_crops, L, heads, out, stats, st);
}

template <int LP>
int ax_launch(const void* qkv, int n_crops, int L, int heads, const float* simmap_t, float sim_weight, void* out, cudaStream_t st) {
  const int width = heads * AT_HD;
  CUtensorMap tkv;
  if (int rc = at_make_map(&tkv, qkv, (long long)n_crops * L, 3 * width, AtCfg<LP>::KV_BOX)) return rc;
  CSEG_SET_SMEM(attention_tc_exp_kernel<LP>, AxCfg<LP>::SMEM);
  const int items = n_crops * heads;
  cseg_launch(attention_tc_exp_kernel<LP>, dim3(std::min(items, sm_count())), dim3(AT_THREADS), AxCfg<LP>::SMEM, st, tkv, L, heads,
              items, (bf16*)out, 0.125f * 1.4426950408889634f, simmap_t, sim_weight);
  CSEG_LAUNCH_CHECK("attention_tc_exp");
  return 0;
}

extern "C" int cseg_attention_experimental_tc(const void* qkv, int n_crops, int L, int heads, const float* simmap_t,
                                              float sim_weight, void* out, void* stream) {
  static_assert(AT_LP == CSEG_SIMT_COLS && 272 == CSEG_SIMT_COLS_MAX,
                "the padded key count is the column count of the transposed similarity map");
  CSEG_REQUIRE(n_crops > 0 && heads > 0 && L >= 17 && L <= CSEG_SIMT_COLS_MAX, "attention_experimental_tc: n_crops=%d heads=%d L=%d",
               n_crops, heads, L);
  CSEG_REQUIRE((((uintptr_t)qkv | (uintptr_t)out) & 15) == 0, "attention_experimental_tc: pointers must be 16-byte aligned");
  if (L <= AT_LP) return ax_launch<AT_LP>(qkv, n_crops, L, heads, simmap_t, sim_weight, out, (cudaStream_t)stream);
  return ax_launch<272>(qkv, n_crops, L, heads, simmap_t, sim_weight, out, (cudaStream_t)stream);
}
