// bf16 GEMM on the 5th-gen tensor cores: TMA (cp.async.bulk.tensor, SWIZZLE_128B) -> shared memory
// -> tcgen05.mma (cta_group::1, kind::f16, M=128) -> fp32 accumulator in TMEM -> tcgen05.ld epilogue.
//
// C[M,N] = residual + alpha * act(A[M,K] . B[N,K]^T + bias)      A, B bf16 K-major (row-major).
//
// Replaces the cuBLAS/cuDNN calls behind nn.Linear / nn.MultiheadAttention in/out projections /
// the 1x1 convolutions of the reference (open_clip/transformer.py:204,211-215,560,770;
// simfeatup_dev/upsamplers.py:218-223,325).
//
// Kernel structure: see the comment above gemm_bf16_tcgen05_kernel (persistent, warp-specialised).
#include "common.cuh"
#include "tc_common.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace {

#ifndef CSEG_EPI_DIRECT
#define CSEG_EPI_DIRECT 0   /* measured slower than the staged epilogue on every pipeline shape */
#endif
constexpr bool EPI_DIRECT = CSEG_EPI_DIRECT != 0;   // row-per-lane epilogue (no smem staging) for full, aligned chunks

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one SWIZZLE_128B row

struct EpiParams {
  const float* bias;
  const void* residual;
  int ldr;
  int res_bf16;
  float alpha;
  int act;
  int out_bf16;
  void* C;
  int ldc;
  int M, N;
  int diag_rows;   // > 0: only tiles that touch the block diagonal (square blocks of diag_rows rows / columns) are computed
  // RES == 3 (compact block-diagonal store, fp32): element (block b, i, j) goes to C[b * cmp_cs + (i - cmp_skip) * cmp_rs +
  // (j - cmp_skip)] for i, j >= cmp_skip; entries that pair rows of different blocks are dropped
  long long cmp_cs;
  int cmp_rs, cmp_skip, cmp_mode;
  // > 0: both operands are [hi | lo] bf16 splits of split_kb k-blocks each and the K loop runs hi.hi + hi.lo + lo.hi
  // (three segments of split_kb k-blocks): fp32-grade products on the bf16 tensor cores
  int split_kb;
};

// Vector form of the staged epilogue for full, aligned chunks WITHOUT a residual: a lane owns FOUR consecutive columns of
// rows r4, r4 + 4, ... (r4 = lane / 8): 8 x LDS.128 from the staging tile and 8 vector stores per chunk instead of 32 + 32
// scalar ones (8 instead of 11 instructions per element in the GELU epilogue of fc1, which paces that GEMM once the CTA
// pair has removed the operand-feed limit).  The fp32-residual variants keep the scalar form: with the residual registers
// on top they spill at the 96-register budget of the 18-warp CTA (measured: out-proj 10 % slower).
template <int ACT, int OUTB>
__device__ __forceinline__ void epi_vec_chunk(const float* stg, int SST, const EpiParams& ep, int rbase, int nrows, int col0, int lane) {
  const int r4 = lane >> 3, c4 = (lane & 7) * 4;
  const float alpha = ep.alpha;
  const float4 bv4 = ep.bias != nullptr ? __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + c4)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int row = it * 4 + r4;
    if (row >= nrows) continue;
    const float4 a = *reinterpret_cast<const float4*>(stg + row * SST + c4);
    float x[4] = {a.x + bv4.x, a.y + bv4.y, a.z + bv4.z, a.w + bv4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (ACT == CSEG_ACT_GELU) x[e] = OUTB ? gelu_tanh(x[e]) : gelu_fast(x[e]);
      else if (ACT == CSEG_ACT_QUICKGELU) x[e] = quick_gelu(x[e]);
      x[e] *= alpha;
    }
    if (OUTB) {
      const __nv_bfloat162 lo = __floats2bfloat162_rn(x[0], x[1]), hi = __floats2bfloat162_rn(x[2], x[3]);
      uint2 u;
      u.x = *reinterpret_cast<const uint32_t*>(&lo);
      u.y = *reinterpret_cast<const uint32_t*>(&hi);
      *reinterpret_cast<uint2*>((bf16*)ep.C + (size_t)(rbase + row) * ep.ldc + col0 + c4) = u;
    } else {
      *reinterpret_cast<float4*>((float*)ep.C + (size_t)(rbase + row) * ep.ldc + col0 + c4) = make_float4(x[0], x[1], x[2], x[3]);
    }
  }
}
__device__ __forceinline__ bool epi_vec_ok(const EpiParams& ep, int col0, bool outb) {
  if (col0 + 32 > ep.N || (ep.ldc & 3) != 0) return false;
  if (((uintptr_t)ep.C & (outb ? 7 : 15)) != 0) return false;
  if (ep.bias != nullptr && ((uintptr_t)ep.bias & 15) != 0) return false;
  return true;
}

// Persistent, warp-specialised kernel.  One CTA per SM loops over output tiles (tile id -> (m, n) with m
// fastest, so concurrently running CTAs share the B tile in L2):
//   warp 0      TMA producer      smem ring of STAGES x (A 128x64 + B BNx64), full/empty mbarriers
//   warp 1      TMEM allocator + MMA issuer; ACC accumulator stages of BN columns in TMEM
//   warps 2..   epilogue: 4 TMEM lane quadrants (= warp id % 4) x column groups; each warp drains one or
//               two 32x32 chunks per tile: tcgen05.ld -> (after its last chunk) release the accumulator
//               stage -> stage through shared memory -> row-coalesced residual loads / C stores.
// Tile widths: BN = 64 / 128 (4 accumulator stages) and 192 / 256 (2 stages).  The L2 -> SM feed saturates
// at roughly 50-60 B/clk/SM with the ring depth that fits in shared memory, so wide tiles (fewer operand bytes
// per flop) are what lifts the large-N GEMMs; BN = 192 gives N = 768 exactly 100 tiles on 148 SMs.
// The epilogue of tile i overlaps the loads and MMAs of tiles i+1.. (ACC up to 4), which is what the
// skinny-K GEMMs of the JBU (K = 128) and the GELU epilogues need: they are epilogue-bound.
template <int BN, int STAGES>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int ACC = (512 / BN) < 4 ? (512 / BN) : 4;   // accumulator stages in TMEM
  static constexpr int TMEM_COLS = (ACC * BN <= 256) ? 256 : 512; // allocation must be a power of two
  static constexpr int NCHUNK = BN / 32;                          // 32-column chunks per tile
  static constexpr int CGROUPS = NCHUNK <= 4 ? NCHUNK : (NCHUNK % 3 == 0 ? 3 : 4);  // column groups of warps
  static constexpr int EPI_WARPS = 4 * CGROUPS;                   // each warp drains NCHUNK / CGROUPS chunks
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
  static constexpr int SST = 36;                                 // staging row stride (floats); 36/4 odd
  static constexpr int STG_OFF = STAGES * STAGE_BYTES;
  static constexpr int STG_BYTES = EPI_WARPS * 32 * SST * 4;
  static constexpr int BAR_OFF = STG_OFF + STG_BYTES;
  static constexpr int NBARS = 2 * STAGES + 2 * ACC;
  static constexpr int SMEM_BYTES = BAR_OFF + NBARS * 8 + 16 + 1024;  // + alignment slack
};


// ACT: CSEG_ACT_*; OUTB: C is bf16; RES: 0 none, 1 fp32 residual, 2 bf16 residual (compile-time so the
// per-element epilogue is ~10 instructions: the skinny-K GEMMs are bound by epilogue instruction issue)
template <int BN, int STAGES, int ACT, int OUTB, int RES>
__global__ void __launch_bounds__(Cfg<BN, STAGES>::THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int K,
                         int m_tiles, int num_tiles, EpiParams ep) {
  using C = Cfg<BN, STAGES>;
  constexpr int ACC = C::ACC;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic on the __shared__ array: accesses compile to LDS / STS, not generic LD / ST
  uint64_t* bars = (uint64_t*)(smem + C::BAR_OFF);
  uint32_t* tmem_slot = (uint32_t*)(bars + C::NBARS);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + STAGES * 8;
  const uint32_t tfull0 = empty0 + STAGES * 8, tempty0 = tfull0 + ACC * 8;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (K + BK - 1) / BK;  // a K tail is zero-filled by TMA (OOB fill) in both operands
  // tile order: the n-tiles of one m-panel run back to back when A is the big operand (m_tiles >= n_tiles:
  // the A panel is fetched from HBM once and re-read from L2), otherwise m fastest (B panel shared).
  const int n_tiles = num_tiles / m_tiles;
  const bool n_fast = m_tiles >= n_tiles;
#define TILE_M(t) (n_fast ? (t) / n_tiles : (t) % m_tiles)
#define TILE_N(t) (n_fast ? (t) % n_tiles : (t) / m_tiles)
  // Tile enumeration, identical in every warp role: v -> (m, n).  Block-diagonal products (per-crop Gram matrices in one
  // launch, ep.diag_rows > 0) only visit the tiles whose row range shares a diagonal block with their column range: for
  // column tile n those are the m-tiles m_lo(n) .. m_hi(n); v = n * DM + dm enumerates them with a fixed bound DM per column,
  // so a CTA evaluates a handful of candidates instead of scanning all m_tiles x n_tiles tile ids.
  const int DM = ep.diag_rows > 0 ? (BN + 2 * ep.diag_rows + BM - 1) / BM + 1 : 1;
  const int n_iter = ep.diag_rows > 0 ? n_tiles * DM : num_tiles;
  auto tile_mn = [&](int v, int& tm, int& tn) -> bool {
    if (ep.diag_rows <= 0) {
      tm = TILE_M(v);
      tn = TILE_N(v);
      return true;
    }
    tn = v / DM;
    const int dm = v - tn * DM, Lb = ep.diag_rows, n0 = tn * BN;
    const int c_lo = n0 / Lb, c_hi = (min(n0 + BN, ep.N) - 1) / Lb;
    const int m_lo = (c_lo * Lb) / BM, m_hi = (min((c_hi + 1) * Lb, ep.M) - 1) / BM;
    tm = m_lo + dm;
    return tm <= m_hi;
  };

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + s * 8, 1);
      mbar_init(empty0 + s * 8, 1);
    }
    for (int s = 0; s < ACC; ++s) {
      mbar_init(tfull0 + s * 8, 1);
      mbar_init(tempty0 + s * 8, C::EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(C::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_grid_sync();   // everything above is input-independent and overlaps the tail of the previous kernel
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_iter; tile += gridDim.x) {
        int tm, tn;
        if (!tile_mn(tile, tm, tn)) continue;
        const int m0 = tm * BM, n0 = tn * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(empty0 + s * 8, ph ^ 1);
          mbar_expect_tx(full0 + s * 8, C::STAGE_BYTES);
          const uint32_t a_dst = smem_base + s * C::STAGE_BYTES;
          int ka = kb, kb_ = kb;
          if (ep.split_kb > 0) {                     // segments: (A hi, B hi), (A hi, B lo), (A lo, B hi)
            const int seg = kb / ep.split_kb, r = kb - seg * ep.split_kb;
            ka = (seg == 2) ? ep.split_kb + r : r;
            kb_ = (seg == 1) ? ep.split_kb + r : r;
          }
          tma_load_2d(a_dst, &tmA, full0 + s * 8, ka * BK, m0);
          tma_load_2d(a_dst + C::A_BYTES, &tmB, full0 + s * 8, kb_ * BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    // the whole warp runs this loop (warp-uniform control flow keeps descriptors and barrier addresses in uniform
    // registers); one elected lane issues the MMAs and the commits
    {
      constexpr uint32_t idesc = make_idesc(BM, BN);
      uint32_t it = 0, tl = 0;
      for (int tile = blockIdx.x; tile < n_iter; tile += gridDim.x) {
        int tm, tn;
        if (!tile_mn(tile, tm, tn)) continue;
        const uint32_t as = tl % ACC, aph = (tl / ACC) & 1;
        ++tl;
        mbar_wait(tempty0 + as * 8, aph ^ 1);      // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(full0 + s * 8, ph);
          tc_fence_after();
          const uint64_t adesc = make_sdesc(smem_base + s * C::STAGE_BYTES);
          const uint64_t bdesc = make_sdesc(smem_base + s * C::STAGE_BYTES + C::A_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // +32 B per K=16 step inside the 128 B swizzle row -> +2 in the (addr >> 4) field
              umma_f16(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(empty0 + s * 8);  // frees the smem stage once these MMAs have read it
            if (kb == num_kb - 1) umma_commit(tfull0 + as * 8);   // accumulator complete
          }
          __syncwarp();
        }
      }
    }
  } else {
    const int ew = warp - 2;
    const int lg = warp & 3;            // TMEM lane quadrant this warp may read (= warp id % 4)
    const int cg = ew >> 2;             // column group: chunks cg, cg + CGROUPS, ...
    constexpr int SST = C::SST, CPW = C::NCHUNK / C::CGROUPS;
    float* stg = reinterpret_cast<float*>(smem + C::STG_OFF) + ew * 32 * SST;
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < n_iter; tile += gridDim.x) {
      int tm, tn;
      if (!tile_mn(tile, tm, tn)) continue;
      const int m0 = tm * BM, n0 = tn * BN;
      const uint32_t as = tl % ACC, aph = (tl / ACC) & 1;
      ++tl;
      const int rbase = m0 + lg * 32;
      const int nrows = min(32, ep.M - rbase);
#pragma unroll
      for (int ci = 0; ci < CPW; ++ci) {
        const int cchunk = cg + ci * C::CGROUPS;
        const int col0 = n0 + cchunk * 32;
        const int col = col0 + lane;
        const bool col_ok = col < ep.N;
        // residual + bias of the chunk are fetched BEFORE the accumulator is read (for the first chunk:
        // before waiting for it), so their latency hides behind the main loop / the previous chunk.
        float res[32];
        const bool direct = EPI_DIRECT && RES != 3 && col0 + 32 <= ep.N && (ep.ldc & 7) == 0 && (RES == 0 || (ep.ldr & 7) == 0);
        if ((RES == 1 || RES == 2) && col_ok && nrows > 0 && !direct) {
          if (RES == 2) {
            const bf16* rp = (const bf16*)ep.residual + (size_t)rbase * ep.ldr + col;
#pragma unroll
            for (int rr = 0; rr < 32; ++rr) res[rr] = (rr < nrows) ? __bfloat162float(rp[(size_t)rr * ep.ldr]) : 0.f;
          } else {
            const float* rp = (const float*)ep.residual + (size_t)rbase * ep.ldr + col;
#pragma unroll
            for (int rr = 0; rr < 32; ++rr) res[rr] = (rr < nrows) ? rp[(size_t)rr * ep.ldr] : 0.f;
          }
        }
        const float bv = (ep.bias != nullptr && col_ok) ? __ldg(ep.bias + col) : 0.f;
        if (ci == 0) {
          mbar_wait(tfull0 + as * 8, aph);
          tc_fence_after();
        }
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * BN + cchunk * 32), r);
        if (ci == CPW - 1) {            // last chunk of this warp: its part of the accumulator is in registers
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty0 + as * 8);
        }
        if (col0 >= ep.N || nrows <= 0) continue;    // warp-uniform
        if (direct) {
          // row-per-lane epilogue straight from the TMEM layout: lane = row, 32 consecutive columns.  Each lane
          // writes one contiguous 64 B (bf16) / 128 B (fp32) segment with 16-byte stores; bias is a broadcast load.
          const int row = rbase + lane;
          if (lane < nrows) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            if (ep.bias != nullptr) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + j));
                v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
              }
            }
            if (ACT == CSEG_ACT_GELU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = OUTB ? gelu_tanh(v[j]) : gelu_fast(v[j]);
            } else if (ACT == CSEG_ACT_QUICKGELU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = quick_gelu(v[j]);
            }
            const float alpha = ep.alpha;
            if (RES == 1) {
              const float4* rp = reinterpret_cast<const float4*>((const float*)ep.residual + (size_t)row * ep.ldr + col0);
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 q4 = rp[j >> 2];
                v[j] = fmaf(v[j], alpha, q4.x); v[j + 1] = fmaf(v[j + 1], alpha, q4.y);
                v[j + 2] = fmaf(v[j + 2], alpha, q4.z); v[j + 3] = fmaf(v[j + 3], alpha, q4.w);
              }
            } else if (RES == 2) {
              const uint4* rp = reinterpret_cast<const uint4*>((const bf16*)ep.residual + (size_t)row * ep.ldr + col0);
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                const uint4 u = rp[j >> 3];
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = __bfloat1622float2(h[e]);
                  v[j + 2 * e] = fmaf(v[j + 2 * e], alpha, f.x);
                  v[j + 2 * e + 1] = fmaf(v[j + 2 * e + 1], alpha, f.y);
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] *= alpha;
            }
            if (OUTB) {
              uint4* cp = reinterpret_cast<uint4*>((bf16*)ep.C + (size_t)row * ep.ldc + col0);
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 pk;
                __nv_bfloat162 t0 = __floats2bfloat162_rn(v[j], v[j + 1]), t1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
                __nv_bfloat162 t2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]), t3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
                pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
                pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
                cp[j >> 3] = pk;
              }
            } else {
              float4* cp = reinterpret_cast<float4*>((float*)ep.C + (size_t)row * ep.ldc + col0);
#pragma unroll
              for (int j = 0; j < 32; j += 4) cp[j >> 2] = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
          }
          continue;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<uint4*>(stg + lane * SST + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
        __syncwarp();
        if (RES == 0 && !direct && epi_vec_ok(ep, col0, OUTB != 0)) {       // warp-uniform
          epi_vec_chunk<ACT, OUTB>(stg, SST, ep, rbase, nrows, col0, lane);
          __syncwarp();
          continue;
        }
        if (RES == 3 && ep.cmp_mode == 1) {
          // row blocks of 32, column-major inside a block ([i / 32][j][i % 32], cmp_rs columns per block): lane = ROW here, so
          // the 32 rows of one column are (up to a block change) consecutive floats -- coalesced stores
          const int Lb = ep.diag_rows, skip = ep.cmp_skip;
          const int row = rbase + lane, blk = row / Lb, i = row - blk * Lb;
          float* cb = (float*)ep.C + (size_t)blk * ep.cmp_cs + (size_t)(i >> 5) * ep.cmp_rs * 32 + (i & 31);
          const int jb = col0 - blk * Lb;
          const bool row_ok = lane < nrows && i >= skip;
          const int jmax = min(32, ep.N - col0);
#pragma unroll 8
          for (int jj = 0; jj < 32; ++jj) {
            const int j = jb + jj;
            if (row_ok && jj < jmax && j >= skip && j < Lb) cb[(size_t)j * 32] = stg[lane * SST + jj] * ep.alpha;
          }
        }
        if (col_ok) {
          const float alpha = ep.alpha;
          auto finish = [&](int rr) -> float {
            float x = stg[rr * SST + lane] + bv;
            if (ACT == CSEG_ACT_GELU) x = OUTB ? gelu_tanh(x) : gelu_fast(x);
            else if (ACT == CSEG_ACT_QUICKGELU) x = quick_gelu(x);
            return (RES == 1 || RES == 2) ? fmaf(x, alpha, res[rr]) : x * alpha;
          };
          if (RES == 3) {
            const int Lb = ep.diag_rows, skip = ep.cmp_skip;
            float* cf = (float*)ep.C;
            if (ep.cmp_mode == 0) {
              int blk = rbase / Lb, i = rbase - blk * Lb;
#pragma unroll 4
              for (int rr = 0; rr < 32; ++rr) {
                if (rr < nrows) {
                  const int j = col - blk * Lb;
                  if (i >= skip && j >= skip && j < Lb)
                    cf[(size_t)blk * ep.cmp_cs + (size_t)(i - skip) * ep.cmp_rs + (j - skip)] = stg[rr * SST + lane] * alpha;
                }
                if (++i == Lb) { i = 0; ++blk; }
              }
            }
          } else if (OUTB) {
            bf16* cp = (bf16*)ep.C + (size_t)rbase * ep.ldc + col;
            if (nrows == 32) {
#pragma unroll
              for (int rr = 0; rr < 32; ++rr) cp[(size_t)rr * ep.ldc] = __float2bfloat16_rn(finish(rr));
            } else {
#pragma unroll
              for (int rr = 0; rr < 32; ++rr)
                if (rr < nrows) cp[(size_t)rr * ep.ldc] = __float2bfloat16_rn(finish(rr));
            }
          } else {
            float* cp = (float*)ep.C + (size_t)rbase * ep.ldc + col;
            if (nrows == 32) {
#pragma unroll
              for (int rr = 0; rr < 32; ++rr) cp[(size_t)rr * ep.ldc] = finish(rr);
            } else {
#pragma unroll
              for (int rr = 0; rr < 32; ++rr)
                if (rr < nrows) cp[(size_t)rr * ep.ldc] = finish(rr);
            }
          }
        }
        __syncwarp();   // staging buffer is reused by this warp's next chunk / tile
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2D bf16 K-major tensor map: dims {K, rows}, box {64, box_rows}, 128B swizzle, zero OOB fill
int make_map(CUtensorMap* m, const void* base, int rows, int K, int ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) CSEG_FAIL(CSEG_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) CSEG_FAIL(CSEG_ECUDA, "cuTensorMapEncodeTiled failed (%d) rows=%d K=%d ld=%d", (int)r, rows, K, ld);
  return 0;
}

template <int BN, int STAGES, int ACT, int OUTB, int RES>
int launch3(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const EpiParams& ep, cudaStream_t st) {
  using C = Cfg<BN, STAGES>;
  CSEG_SET_SMEM((gemm_bf16_tcgen05_kernel<BN, STAGES, ACT, OUTB, RES>), C::SMEM_BYTES);
  const int m_tiles = cdiv(M, BM), num_tiles = m_tiles * cdiv(N, BN);
  const int grid = std::min(num_tiles, sm_count());
  cseg_launch(gemm_bf16_tcgen05_kernel<BN, STAGES, ACT, OUTB, RES>, dim3(grid), dim3(C::THREADS), C::SMEM_BYTES, st, ta, tb, K, m_tiles,
                                                                                              num_tiles, ep);
  CSEG_LAUNCH_CHECK("gemm_bf16_tcgen05");
  return 0;
}
template <int BN, int STAGES, int ACT, int OUTB>
int launch2(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const EpiParams& ep, cudaStream_t st) {
  if (ep.residual == nullptr) return launch3<BN, STAGES, ACT, OUTB, 0>(ta, tb, M, N, K, ep, st);
  if (ep.res_bf16) return launch3<BN, STAGES, ACT, OUTB, 2>(ta, tb, M, N, K, ep, st);
  return launch3<BN, STAGES, ACT, OUTB, 1>(ta, tb, M, N, K, ep, st);
}
template <int BN, int STAGES, int ACT>
int launch1(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const EpiParams& ep, cudaStream_t st) {
  if (ep.out_bf16) return launch2<BN, STAGES, ACT, 1>(ta, tb, M, N, K, ep, st);
  return launch2<BN, STAGES, ACT, 0>(ta, tb, M, N, K, ep, st);
}
template <int BN, int STAGES>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const EpiParams& ep, cudaStream_t st) {
  if (ep.act == CSEG_ACT_GELU) return launch1<BN, STAGES, CSEG_ACT_GELU>(ta, tb, M, N, K, ep, st);
  if (ep.act == CSEG_ACT_QUICKGELU) return launch1<BN, STAGES, CSEG_ACT_QUICKGELU>(ta, tb, M, N, K, ep, st);
  return launch1<BN, STAGES, CSEG_ACT_NONE>(ta, tb, M, N, K, ep, st);
}

// =====================================================================================================
// CTA-pair variant (cta_group::2) for the big ViT GEMMs: two CTAs of a cluster (the two SMs of a TPC) compute ONE
// 256 x BN output tile.  Each CTA stages its own 128 rows of A and HALF of the B tile (BN / 2 rows) per k-block -- 32 KB
// per stage instead of 48 KB: the single-CTA kernel is bound by the L2 -> SM operand feed on these shapes -- and the
// leader CTA (cluster rank 0) issues tcgen05.mma.cta_group::2 (M = 256), which reads both CTAs' shared memory and writes
// each CTA's 128 accumulator rows into its own TMEM.
//   full[s]    lives in the leader only: armed by the leader's producer with the bytes of BOTH CTAs; every TMA load
//              (cp.async.bulk.tensor ... cta_group::2) signals the leader's barrier (peer bit of the address cleared)
//   empty[s]   one per CTA: the leader's tcgen05.commit is multicast to both
//   tfull[a]   one per CTA (multicast commit); tempty[a] lives in the leader and counts the epilogue warps of both CTAs
//              (the peer arrives remotely)
// The epilogue is the staged one of the single-CTA kernel (RES 0 / 1, fp32 or bf16 output, optional activation).
// =====================================================================================================
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;       // clears the CTA-pair bit of a shared::cluster address: the even (leader) CTA

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar) {       // arrives on `bar` in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((unsigned short)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {    // arrive on the leader CTA's barrier at this offset
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & PEER_MASK) : "memory");
}

template <int BN, int STAGES>
struct Cfg2 {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (BN / 2) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int ACC = 512 / BN;
  static constexpr int NCHUNK = BN / 32, CGROUPS = 4, EPI_WARPS = 16, THREADS = 64 + 32 * EPI_WARPS;
  static constexpr int SST = 36;
  static constexpr int STG_OFF = STAGES * STAGE_BYTES;
  static constexpr int STG_BYTES = EPI_WARPS * 32 * SST * 4;
  static constexpr int BAR_OFF = STG_OFF + STG_BYTES;
  static constexpr int NBARS = 2 * STAGES + 2 * ACC;
  static constexpr int SMEM_BYTES = BAR_OFF + NBARS * 8 + 16 + 1024;
};

template <int BN, int STAGES, int ACT, int OUTB, int RES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Cfg2<BN, STAGES>::THREADS, 1)
gemm_bf16_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int K,
                              int m2_tiles, int num_tiles, EpiParams ep) {
  using C = Cfg2<BN, STAGES>;
  constexpr int ACC = C::ACC;
  static_assert(RES == 0 || RES == 1, "the pair kernel serves the ViT GEMMs: no residual or an fp32 residual");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = (uint64_t*)(smem + C::BAR_OFF);
  uint32_t* tmem_slot = (uint32_t*)(bars + C::NBARS);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + STAGES * 8;
  const uint32_t tfull0 = empty0 + STAGES * 8, tempty0 = tfull0 + ACC * 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int num_kb = (K + BK - 1) / BK;
  const int n_tiles = num_tiles / m2_tiles;
  const bool n_fast = m2_tiles >= n_tiles;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + s * 8, 1);
      mbar_init(empty0 + s * 8, 1);
    }
    for (int s = 0; s < ACC; ++s) {
      mbar_init(tfull0 + s * 8, 1);
      mbar_init(tempty0 + s * 8, 2 * C::EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync_all();            // the peer's barriers exist before anything arrives on them remotely
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  pdl_grid_sync();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = pair; tile < num_tiles; tile += n_pairs) {
        const int tm = n_fast ? tile / n_tiles : tile % m2_tiles, tn = n_fast ? tile % n_tiles : tile / m2_tiles;
        const int m0 = tm * 2 * BM + (int)rank * BM, n0 = tn * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(empty0 + s * 8, ph ^ 1);
          if (leader) mbar_expect_tx(full0 + s * 8, 2 * C::STAGE_BYTES);
          const uint32_t a_dst = smem_base + s * C::STAGE_BYTES;
          tma_load_2d_2sm(a_dst, &tmA, full0 + s * 8, kb * BK, m0);
          tma_load_2d_2sm(a_dst + C::A_BYTES, &tmB, full0 + s * 8, kb * BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      constexpr uint32_t idesc = make_idesc(2 * BM, BN);
      uint32_t it = 0, tl = 0;
      for (int tile = pair; tile < num_tiles; tile += n_pairs, ++tl) {
        const uint32_t as = tl % ACC, aph = (tl / ACC) & 1;
        mbar_wait(tempty0 + as * 8, aph ^ 1);      // the epilogue warps of BOTH CTAs have drained this accumulator stage
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(full0 + s * 8, ph);
          tc_fence_after();
          const uint64_t adesc = make_sdesc(smem_base + s * C::STAGE_BYTES);
          const uint64_t bdesc = make_sdesc(smem_base + s * C::STAGE_BYTES + C::A_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma2_f16(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma2_commit_mc(empty0 + s * 8);
            if (kb == num_kb - 1) umma2_commit_mc(tfull0 + as * 8);
          }
          __syncwarp();
        }
      }
    }
  } else {
    const int ew = warp - 2;
    const int lg = warp & 3;
    const int cg = ew >> 2;
    constexpr int SST = C::SST, CPW = C::NCHUNK / C::CGROUPS;
    float* stg = reinterpret_cast<float*>(smem + C::STG_OFF) + ew * 32 * SST;
    uint32_t tl = 0;
    for (int tile = pair; tile < num_tiles; tile += n_pairs, ++tl) {
      const int tm = n_fast ? tile / n_tiles : tile % m2_tiles, tn = n_fast ? tile % n_tiles : tile / m2_tiles;
      const int m0 = tm * 2 * BM + (int)rank * BM, n0 = tn * BN;
      const uint32_t as = tl % ACC, aph = (tl / ACC) & 1;
      const int rbase = m0 + lg * 32;
      const int nrows = min(32, ep.M - rbase);
#pragma unroll
      for (int ci = 0; ci < CPW; ++ci) {
        const int cchunk = cg + ci * C::CGROUPS;
        const int col0 = n0 + cchunk * 32;
        const int col = col0 + lane;
        const bool col_ok = col < ep.N;
        float res[32];
        if (RES == 1 && col_ok && nrows > 0) {
          const float* rp = (const float*)ep.residual + (size_t)rbase * ep.ldr + col;
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) res[rr] = (rr < nrows) ? rp[(size_t)rr * ep.ldr] : 0.f;
        }
        const float bv = (ep.bias != nullptr && col_ok) ? __ldg(ep.bias + col) : 0.f;
        if (ci == 0) {
          mbar_wait(tfull0 + as * 8, aph);
          tc_fence_after();
        }
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * BN + cchunk * 32), r);
        if (ci == CPW - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(tempty0 + as * 8);
        }
        if (col0 >= ep.N || nrows <= 0) continue;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<uint4*>(stg + lane * SST + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
        __syncwarp();
        if (RES == 0 && epi_vec_ok(ep, col0, OUTB != 0)) {       // warp-uniform
          epi_vec_chunk<ACT, OUTB>(stg, SST, ep, rbase, nrows, col0, lane);
          __syncwarp();
          continue;
        }
        if (col_ok) {
          const float alpha = ep.alpha;
          auto finish = [&](int rr) -> float {
            float x = stg[rr * SST + lane] + bv;
            if (ACT == CSEG_ACT_GELU) x = OUTB ? gelu_tanh(x) : gelu_fast(x);
            else if (ACT == CSEG_ACT_QUICKGELU) x = quick_gelu(x);
            return RES == 1 ? fmaf(x, alpha, res[rr]) : x * alpha;
          };
          if (OUTB) {
            bf16* cp = (bf16*)ep.C + (size_t)rbase * ep.ldc + col;
#pragma unroll
            for (int rr = 0; rr < 32; ++rr)
              if (rr < nrows) cp[(size_t)rr * ep.ldc] = __float2bfloat16_rn(finish(rr));
          } else {
            float* cp = (float*)ep.C + (size_t)rbase * ep.ldc + col;
#pragma unroll
            for (int rr = 0; rr < 32; ++rr)
              if (rr < nrows) cp[(size_t)rr * ep.ldc] = finish(rr);
          }
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();            // neither CTA leaves (or frees TMEM) while the other may still signal it
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// =====================================================================================================
// Fused final 1x1 conv + L2-normalise + cosine logits (K12 + K13 of SURVEY 2.2):
//   out[p, :] = y[p, :] + alpha * (y[p, :] . W^T + b)        (simfeatup_dev/upsamplers.py:325)
//   logits[crop, q, pix] = <out[p] / |out[p]|, T[q]> (+ cls bias)   (segmentor.py:374-375,378-379)
// The C x 224^2 feature map `out` never reaches HBM: the GEMM epilogue accumulates |out|^2 and the Q dot
// products per pixel.  Same persistent TMA / tcgen05 / TMEM pipeline as the generic kernel, but the unit of
// scheduling is an M-panel: the NT = C/128 n-tiles of a panel go to the same CTA (one TMEM accumulator stage
// each), so the per-row partial sums stay in registers across n-tiles.  Epilogue warps use the native TMEM
// layout (lane = row), which makes the per-row reductions thread-local.
// =====================================================================================================
template <int QT>
struct NsCfg {
  static constexpr int BN = 128, STAGES = (QT <= 8 ? 4 : 3), NT_MAX = 4;
  static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_WARPS = 16, THREADS = 64 + 32 * EPI_WARPS;
  static constexpr int TXT_OFF = STAGES * STAGE_BYTES;               // text  [C_MAX][QT] fp32
  static constexpr int C_MAX = 512;
  static constexpr int TXT_BYTES = C_MAX * QT * 4;
  static constexpr int BIAS_OFF = TXT_OFF + TXT_BYTES;               // bias  [C_MAX] fp32
  static constexpr int PART_OFF = BIAS_OFF + C_MAX * 4;              // part  [2][4 chunks][128 rows][QT+1]
  static constexpr int PART_BYTES = 2 * 4 * 128 * (QT + 1) * 4;
  static constexpr int BAR_OFF = PART_OFF + PART_BYTES;
  static constexpr int NBARS = 2 * STAGES + 2 * NT_MAX;
  static constexpr int SMEM_BYTES = BAR_OFF + NBARS * 8 + 16 + 1024;
};

template <int QT>
__global__ void __launch_bounds__(NsCfg<QT>::THREADS, 1)
gemm_normsim_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int C,
                    const bf16* __restrict__ y, int ldy, const float* __restrict__ bias, float alpha,
                    const float* __restrict__ text, int Q, const float* __restrict__ cls_bias, int hw,
                    float* __restrict__ logits) {
  using Cf = NsCfg<QT>;
  constexpr int BN = Cf::BN, STAGES = Cf::STAGES, ACC = Cf::NT_MAX;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic on the __shared__ array: accesses compile to LDS / STS, not generic LD / ST
  uint64_t* bars = (uint64_t*)(smem + Cf::BAR_OFF);
  uint32_t* tmem_slot = (uint32_t*)(bars + Cf::NBARS);
  float* txt = reinterpret_cast<float*>(smem + Cf::TXT_OFF);
  float* bs = reinterpret_cast<float*>(smem + Cf::BIAS_OFF);
  float* part = reinterpret_cast<float*>(smem + Cf::PART_OFF);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + STAGES * 8;
  const uint32_t tfull0 = empty0 + STAGES * 8, tempty0 = tfull0 + ACC * 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = C / BK, NT = C / BN, panels = (M + BM - 1) / BM;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + s * 8, 1);
      mbar_init(empty0 + s * 8, 1);
    }
    for (int s = 0; s < ACC; ++s) {
      mbar_init(tfull0 + s * 8, 1);
      mbar_init(tempty0 + s * 8, Cf::EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_grid_sync();   // everything above is input-independent and overlaps the tail of the previous kernel
  // text (transposed to [c][QT], zero padded) and bias into shared memory
  for (int e = threadIdx.x; e < C * QT; e += Cf::THREADS) {
    const int c = e / QT, q = e % QT;
    txt[e] = (q < Q) ? text[(size_t)q * C + c] : 0.f;
  }
  for (int c = threadIdx.x; c < C; c += Cf::THREADS) bs[c] = bias ? bias[c] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int panel = blockIdx.x; panel < panels; panel += gridDim.x)
        for (int nt = 0; nt < NT; ++nt)
          for (int kb = 0; kb < num_kb; ++kb, ++it) {
            const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
            mbar_wait(empty0 + s * 8, ph ^ 1);
            mbar_expect_tx(full0 + s * 8, Cf::STAGE_BYTES);
            const uint32_t a_dst = smem_base + s * Cf::STAGE_BYTES;
            tma_load_2d(a_dst, &tmA, full0 + s * 8, kb * BK, panel * BM);
            tma_load_2d(a_dst + Cf::A_BYTES, &tmB, full0 + s * 8, kb * BK, nt * BN);
          }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN);
      uint32_t it = 0, tl = 0;
      for (int panel = blockIdx.x; panel < panels; panel += gridDim.x)
        for (int nt = 0; nt < NT; ++nt, ++tl) {
          const uint32_t as = tl % ACC, aph = (tl / ACC) & 1;
          mbar_wait(tempty0 + as * 8, aph ^ 1);
          tc_fence_after();
          const uint32_t tacc = tmem_base + as * BN;
          for (int kb = 0; kb < num_kb; ++kb, ++it) {
            const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
            mbar_wait(full0 + s * 8, ph);
            tc_fence_after();
            const uint64_t adesc = make_sdesc(smem_base + s * Cf::STAGE_BYTES);
            const uint64_t bdesc = make_sdesc(smem_base + s * Cf::STAGE_BYTES + Cf::A_BYTES);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_f16(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit(empty0 + s * 8);
          }
          umma_commit(tfull0 + as * 8);
        }
    }
  } else {
    const int ew = warp - 2, lg = warp & 3, cchunk = ew >> 2;
    uint32_t tl = 0, pi = 0;
    for (int panel = blockIdx.x; panel < panels; panel += gridDim.x, ++pi) {
      const int row = panel * BM + lg * 32 + lane;
      const bool row_ok = row < M;
      float ss = 0.f, dot[QT];
#pragma unroll
      for (int q = 0; q < QT; ++q) dot[q] = 0.f;
      for (int nt = 0; nt < NT; ++nt, ++tl) {
        const uint32_t as = tl % ACC, aph = (tl / ACC) & 1;
        const int col0 = nt * BN + cchunk * 32;
        // residual y[row, col0 .. col0+31] (64 B per lane), prefetched before the accumulator is ready
        uint4 yv[4];
#pragma unroll
        for (int v = 0; v < 4; ++v)
          yv[v] = row_ok ? *reinterpret_cast<const uint4*>(y + (size_t)row * ldy + col0 + v * 8) : make_uint4(0, 0, 0, 0);
        mbar_wait(tfull0 + as * 8, aph);
        tc_fence_after();
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * BN + cchunk * 32), r);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty0 + as * 8);
        const __nv_bfloat162* yh = reinterpret_cast<const __nv_bfloat162*>(yv);
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float2 yr = __bfloat1622float2(yh[j >> 1]);
          const float o0 = fmaf(alpha, __uint_as_float(r[j]) + bs[col0 + j], yr.x);
          const float o1 = fmaf(alpha, __uint_as_float(r[j + 1]) + bs[col0 + j + 1], yr.y);
          ss = fmaf(o0, o0, ss);
          ss = fmaf(o1, o1, ss);
          const float4* t0 = reinterpret_cast<const float4*>(txt + (col0 + j) * QT);
          const float4* t1 = reinterpret_cast<const float4*>(txt + (col0 + j + 1) * QT);
#pragma unroll
          for (int q4 = 0; q4 < QT / 4; ++q4) {
            const float4 a = t0[q4], b = t1[q4];
            dot[q4 * 4 + 0] = fmaf(o0, a.x, dot[q4 * 4 + 0]); dot[q4 * 4 + 1] = fmaf(o0, a.y, dot[q4 * 4 + 1]);
            dot[q4 * 4 + 2] = fmaf(o0, a.z, dot[q4 * 4 + 2]); dot[q4 * 4 + 3] = fmaf(o0, a.w, dot[q4 * 4 + 3]);
            dot[q4 * 4 + 0] = fmaf(o1, b.x, dot[q4 * 4 + 0]); dot[q4 * 4 + 1] = fmaf(o1, b.y, dot[q4 * 4 + 1]);
            dot[q4 * 4 + 2] = fmaf(o1, b.z, dot[q4 * 4 + 2]); dot[q4 * 4 + 3] = fmaf(o1, b.w, dot[q4 * 4 + 3]);
          }
        }
      }
      // combine the four column chunks of every row in a fixed order (deterministic), then finalise
      float* pb = part + (size_t)(pi & 1) * 4 * 128 * (QT + 1);
      float* mine = pb + ((size_t)cchunk * 128 + lg * 32 + lane) * (QT + 1);
      mine[0] = ss;
#pragma unroll
      for (int q = 0; q < QT; ++q) mine[1 + q] = dot[q];
      asm volatile("bar.sync 1, %0;" ::"n"(32 * NsCfg<QT>::EPI_WARPS) : "memory");   // epilogue warps only
      if (cchunk == 0 && row_ok) {
        float tot[QT + 1];
#pragma unroll
        for (int q = 0; q <= QT; ++q) tot[q] = 0.f;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const float* pp = pb + ((size_t)ch * 128 + lg * 32 + lane) * (QT + 1);
#pragma unroll
          for (int q = 0; q <= QT; ++q) tot[q] += pp[q];
        }
        const float inv = 1.0f / sqrtf(tot[0]);
        const long long crop = row / hw, pix = row % hw;
#pragma unroll
        for (int q = 0; q < QT; ++q)
          if (q < Q) {
            float v = tot[1 + q] * inv;
            if (cls_bias) v += cls_bias[crop * Q + q];
            logits[(crop * Q + q) * hw + pix] = v;
          }
      }
      // the part buffer alternates per panel; a buffer is rewritten two panels later, after another bar.sync
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

template <int QT>
int launch_normsim(const CUtensorMap& ta, const CUtensorMap& tb, int M, int C, const bf16* y, int ldy, const float* bias,
                   float alpha, const float* text, int Q, const float* cls_bias, int hw, float* logits, cudaStream_t st) {
  using Cf = NsCfg<QT>;
  CSEG_SET_SMEM(gemm_normsim_kernel<QT>, Cf::SMEM_BYTES);
  const int panels = cdiv(M, BM);
  cseg_launch(gemm_normsim_kernel<QT>, dim3(std::min(panels, sm_count())), dim3(Cf::THREADS), Cf::SMEM_BYTES, st, 
      ta, tb, M, C, y, ldy, bias, alpha, text, Q, cls_bias, hw, logits);
  CSEG_LAUNCH_CHECK("gemm_normsim");
  return 0;
}


// =====================================================================================================
// Cosine logits from basis coefficients (the "basis path" of the JBU head).
// The JBU stack is linear in its source and acts on every channel alike, so upsampling the identity (one
// channel per low-resolution token) yields coefficients s[p, k] with
//   feature[p] = sum_k s[p, k] * g[k] + b            g = per-crop token features after the final 1x1 conv
// and the normalise + similarity of segmentor.py:374-379 becomes, per pixel p of a crop,
//   |feature|^2 = s^T (g g^T) s + 2 s.(g b) + b.b     <feature, t_q> = s.(g t_q) + b.t_q
// i.e. one GEMM  D = S . [Gram | Aux]^T  with N = T + 16 columns instead of C, whose epilogue needs one
// multiply-add per Gram column (sum_j D[p, j] s[p, j]).  K = T (tokens per crop, <= 256) instead of C.
// Same TMA / tcgen05 / TMEM pipeline as above; one M-panel (128 pixels of one crop) per accumulator stage.
// =====================================================================================================
struct BlCfg {
  static constexpr int STAGES = 5, N_MAX = 256, N2 = 16, KB_MAX = 4;   // N = Gram columns + N2 aux columns, K <= 256
  static constexpr int A_BYTES = BM * BK * 2, B_BYTES = N_MAX * BK * 2; // A ring stage 16 KB; one B k-block 32 KB
  static constexpr int B_OFF = STAGES * A_BYTES;                        // B of the current crop stays resident
  static constexpr int EPI_WARPS = 16, THREADS = 64 + 32 * EPI_WARPS;
  static constexpr int PART_OFF = B_OFF + KB_MAX * B_BYTES;             // part [2][4 column groups][128 rows]
  static constexpr int PART_BYTES = 2 * 4 * 128 * 4;
  static constexpr int BAR_OFF = PART_OFF + PART_BYTES;
  static constexpr int NBARS = 2 * STAGES + 6;
  static constexpr int SMEM_BYTES = BAR_OFF + NBARS * 8 + 16 + 1024;
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

// KS = ceil(T / 16) K-steps (and Gram column groups of 16); panels never straddle crops (hw % 128 == 0).
// gram is the full (n_crops*tstride)^2 product g g^T; the tile of a crop is its diagonal block (rows and columns
// crop*tstride ..; rows / columns past the crop's T tokens meet zero coefficients and never contribute).
// The B tile of a k-block is the Gram rows followed by the 16 aux rows (two TMA loads into adjacent 8-row groups
// of the SWIZZLE_128B layout), so one MMA of N = 16 KS + 16 <= 256 produces both; two accumulator stages.
// Every CTA takes a contiguous range of panels; the B tile of the current crop (all k-blocks, <= 128 KB) stays
// resident in shared memory and only the coefficient panels stream through the TMA ring.
__global__ void __launch_bounds__(BlCfg::THREADS, 1)
basis_logits_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB1,
                    const __grid_constant__ CUtensorMap tmB2, int panels, int hw, int tstride, int KS,
                    const bf16* __restrict__ s, int lds, const float* __restrict__ consts, int Q,
                    const float* __restrict__ cls_bias, float* __restrict__ logits, int n2) {
  using Cf = BlCfg;
  constexpr int STAGES = Cf::STAGES;
  constexpr uint32_t ACC = 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic on the __shared__ array: accesses compile to LDS / STS, not generic LD / ST
  uint64_t* bars = (uint64_t*)(smem + Cf::BAR_OFF);
  uint32_t* tmem_slot = (uint32_t*)(bars + Cf::NBARS);
  float* part = reinterpret_cast<float*>(smem + Cf::PART_OFF);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + STAGES * 8;
  const uint32_t tfull0 = empty0 + STAGES * 8, tempty0 = tfull0 + ACC * 8;
  const uint32_t bfull = tempty0 + ACC * 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N1 = KS * 16, num_kb = (KS + 3) / 4;
  const int panels_per_crop = hw / BM;
  // contiguous panel range per CTA: consecutive panels share the crop, whose Gram / aux tile stays in shared memory
  const int ppc = (panels + gridDim.x - 1) / gridDim.x;
  const int p_begin = blockIdx.x * ppc, p_end = min(panels, p_begin + ppc);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB2) : "memory");
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full0 + i * 8, 1);
      mbar_init(empty0 + i * 8, 1 + Cf::EPI_WARPS);    // the MMA commit and every epilogue warp (reads its coefficients)
    }
    for (int i = 0; i < (int)ACC; ++i) {
      mbar_init(tfull0 + i * 8, 1);
      mbar_init(tempty0 + i * 8, Cf::EPI_WARPS);
    }
    mbar_init(bfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_grid_sync();   // everything above is input-independent and overlaps the tail of the previous kernel
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t b1_bytes = (uint32_t)(N1 * BK * 2);
      const uint32_t b_tx = (uint32_t)num_kb * (b1_bytes + (uint32_t)(n2 * BK * 2));
      uint32_t it = 0;
      int cur_crop = -1;
      for (int panel = p_begin; panel < p_end; ++panel) {
        const int crop = panel / panels_per_crop;
        if (crop != cur_crop) {                       // new crop: reload the resident B once the MMAs that read the old one are done
          // the commit that frees the most recently filled A stage covers every MMA issued before it; the producer
          // has observed all earlier phases of that barrier, so the parity wait cannot alias
          if (it > 0) mbar_wait(empty0 + ((it - 1) % STAGES) * 8, ((it - 1) / STAGES) & 1);
          mbar_expect_tx(bfull, b_tx);
          for (int kb = 0; kb < num_kb; ++kb) {
            const uint32_t b_dst = smem_base + Cf::B_OFF + kb * Cf::B_BYTES;
            tma_load_2d(b_dst, &tmB1, bfull, crop * tstride + kb * BK, crop * tstride);
            tma_load_2d(b_dst + b1_bytes, &tmB2, bfull, crop * tstride + kb * BK, 0);
          }
          cur_crop = crop;
        }
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t st = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(empty0 + st * 8, ph ^ 1);
          mbar_expect_tx(full0 + st * 8, Cf::A_BYTES);
          tma_load_2d(smem_base + st * Cf::A_BYTES, &tmA, full0 + st * 8, kb * BK, panel * BM);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BM, N1 + n2);
      uint32_t it = 0, tl = 0, nb = 0;
      int cur_crop = -1;
      for (int panel = p_begin; panel < p_end; ++panel, ++tl) {
        const uint32_t as = tl % ACC, aph = (tl / ACC) & 1;
        const int crop = panel / panels_per_crop;
        if (crop != cur_crop) {
          mbar_wait(bfull, nb & 1);
          cur_crop = crop;
          ++nb;
        }
        mbar_wait(tempty0 + as * 8, aph ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * 256;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t st = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(full0 + st * 8, ph);
          tc_fence_after();
          const uint64_t adesc = make_sdesc(smem_base + st * Cf::A_BYTES);
          const uint64_t bdesc = make_sdesc(smem_base + Cf::B_OFF + kb * Cf::B_BYTES);
          const int ksteps = min(4, KS - kb * 4);
          for (int k = 0; k < ksteps; ++k)
            umma_f16(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(empty0 + st * 8);
        }
        umma_commit(tfull0 + as * 8);
      }
    }
  } else {
    const int ew = warp - 2, lg = warp & 3, cg = ew >> 2;
    uint32_t tl = 0, it = 0;
    const int r128 = lg * 32 + lane;
    for (int panel = p_begin; panel < p_end; ++panel, ++tl) {
      const uint32_t as = tl % ACC, aph = (tl / ACC) & 1;
      const size_t row = (size_t)panel * BM + r128;
      // this lane's coefficients of the Gram column groups cg, cg+4, ... (32 B each) come from the TMA tiles in shared
      // memory: group cg + 4 i lies in k-block i, chunks 2 cg and 2 cg + 1 of the row (SWIZZLE_128B: chunk ^ (row & 7)).
      // A row-per-lane global load would cost 32 wavefronts per instruction; these are conflict free.
      uint4 sv[4][2];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (i < num_kb) {                                  // warp-uniform; every epilogue warp consumes every k-block
          const uint32_t st = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(full0 + st * 8, ph);
          if (cg + 4 * i < KS) {
            const uint8_t* arow = smem + st * Cf::A_BYTES + r128 * 128;
            sv[i][0] = *reinterpret_cast<const uint4*>(arow + (((2 * cg) ^ (r128 & 7)) << 4));
            sv[i][1] = *reinterpret_cast<const uint4*>(arow + (((2 * cg + 1) ^ (r128 & 7)) << 4));
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(empty0 + st * 8);
          ++it;
        }
      }
      mbar_wait(tfull0 + as * 8, aph);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(lg * 32) << 16) + as * 256;
      float den = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int g = cg + 4 * i;
        if (g < KS) {                                      // warp-uniform
          uint32_t r[16];
          __syncwarp();
          tmem_ld16(trow + (uint32_t)(g * 16), r);
          const __nv_bfloat162* sh = reinterpret_cast<const __nv_bfloat162*>(&sv[i][0]);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 f = __bfloat1622float2(sh[j]);
            den = fmaf(__uint_as_float(r[2 * j]), f.x, den);
            den = fmaf(__uint_as_float(r[2 * j + 1]), f.y, den);
          }
        }
      }
      uint32_t nr[16], nr2[16];                            // aux columns 0..15 (and 16..31 when Q + 1 > 16)
      if (cg == 3) {
        __syncwarp();
        tmem_ld16(trow + (uint32_t)N1, nr);
        if (n2 > 16) {
          __syncwarp();
          tmem_ld16(trow + (uint32_t)N1 + 16, nr2);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + as * 8);
      float* pb = part + (size_t)(tl & 1) * 4 * 128;
      pb[cg * 128 + lg * 32 + lane] = den;
      asm volatile("bar.sync 1, %0;" ::"n"(32 * BlCfg::EPI_WARPS) : "memory");   // epilogue warps only
      if (cg == 3) {
        const int r128 = lg * 32 + lane;
        float hb = 0.f;                                    // aux column Q = s . (g b); select keeps nr[] in registers
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          if (q == Q) hb = __uint_as_float(nr[q]);
          if (n2 > 16 && q + 16 == Q) hb = __uint_as_float(nr2[q]);
        }
        const float d2 = pb[r128] + pb[128 + r128] + pb[256 + r128] + pb[384 + r128] + 2.f * hb + __ldg(consts + Q);
        const float inv = 1.0f / sqrtf(fmaxf(d2, 1e-30f));
        const size_t crop = row / hw, pix = row % hw;
#pragma unroll
        for (int q = 0; q < 16; ++q)
          if (q < Q) {
            float v = (__uint_as_float(nr[q]) + __ldg(consts + q)) * inv;
            if (cls_bias) v += cls_bias[crop * Q + q];
            logits[(crop * Q + q) * hw + pix] = v;
          }
        if (n2 > 16) {
#pragma unroll
          for (int q = 0; q < 15; ++q)
            if (q + 16 < Q) {
              float v = (__uint_as_float(nr2[q]) + __ldg(consts + q + 16)) * inv;
              if (cls_bias) v += cls_bias[crop * Q + q + 16];
              logits[(crop * Q + q + 16) * hw + pix] = v;
            }
        }
      }
      // the part buffer alternates per panel; a buffer is rewritten two panels later, after another bar.sync
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}


// =====================================================================================================
// JBU kernel fix-up in one pass (simfeatup_dev/upsamplers.py:218-223,258-262):
//   out[p, :] = k[p, :] + 0.1 * (W3 . gelu(W0 . k[p, :] + b0) + b3)            k = range x spatial kernel, LDK wide
// Two chained tcgen05 GEMMs per 128-row panel.  The hidden activations never leave the SM: the first epilogue
// writes gelu(.) as bf16 straight into the SWIZZLE_128B K-major layout the second MMA reads; the residual is read
// back from the TMA-loaded k tile in shared memory.  HBM traffic = read k once + write out once.
//   warp 0 TMA producer (W0, W3 once; k panels, 3-slot ring)     warp 1 MMA issuer (MMA1 of panel i+1 before
//   MMA2 of panel i)     warps 2-9 epilogue 1 (D1 -> gelu -> H tile)     warps 10-17 epilogue 2 (D2 + k + b3 -> out);
//   both epilogues: lane quadrant = warp % 4, two warps per quadrant split the column chunks
// =====================================================================================================
template <int LDK>
struct FxCfg {
  static constexpr int NKB = LDK / BK;
  static constexpr int W_BYTES = LDK * LDK * 2, P_BYTES = BM * LDK * 2, KB_W = LDK * 128, KB_P = BM * 128;
  static constexpr int NA = 3;                                      // k-panel ring: one stage more than accumulator stages
  static constexpr int W0_OFF = 0, W3_OFF = W_BYTES, A_OFF = 2 * W_BYTES, H_OFF = A_OFF + NA * P_BYTES;
  static constexpr int BIAS_OFF = H_OFF + 2 * P_BYTES;             // b0[LDK], b3[LDK] fp32
  static constexpr int BAR_OFF = BIAS_OFF + 2 * LDK * 4;
  static constexpr int NBARS = 19;
  static constexpr int SMEM_BYTES = BAR_OFF + NBARS * 8 + 16 + 1024;
  // 8 + 8 epilogue warps: with one warp per scheduler the two epilogues were latency bound (source-level samples: issuing
  // 33 % of the cycles, the rest fixed-latency / shared-memory waits of a single dependent chain)
  static constexpr int E1W = 8, E2W = 8, THREADS = 32 * (2 + E1W + E2W), TMEM_COLS = 4 * LDK;
};

template <int LDK>
__global__ void __launch_bounds__(FxCfg<LDK>::THREADS, 1)
kernel_fixup_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW0,
                    const __grid_constant__ CUtensorMap tmW3, const __grid_constant__ CUtensorMap tmOut, int M,
                    const float* __restrict__ b0, const float* __restrict__ b3) {
  using Cf = FxCfg<LDK>;
  constexpr int NKB = Cf::NKB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic on the __shared__ array: accesses compile to LDS / STS, not generic LD / ST
  uint64_t* bars = (uint64_t*)(smem + Cf::BAR_OFF);
  uint32_t* tmem_slot = (uint32_t*)(bars + Cf::NBARS);
  float* bs0 = reinterpret_cast<float*>(smem + Cf::BIAS_OFF);
  float* bs3 = bs0 + LDK;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t w_full = smem_u32(bars);
  const uint32_t a_full0 = w_full + 8, a_empty0 = a_full0 + 24, d1_full0 = a_empty0 + 24, d1_empty0 = d1_full0 + 16;
  const uint32_t h_full0 = d1_empty0 + 16, h_empty0 = h_full0 + 16, d2_full0 = h_empty0 + 16, d2_empty0 = d2_full0 + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int panels = (M + BM - 1) / BM;
  const int n_my = (panels > (int)blockIdx.x) ? (panels - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW3) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmOut) : "memory");
    mbar_init(w_full, 1);
    for (int i = 0; i < Cf::NA; ++i) {
      mbar_init(a_full0 + i * 8, 1);
      mbar_init(a_empty0 + i * 8, 1);                  // the thread that issued the TMA store of the slot
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(d1_full0 + i * 8, 1);
      mbar_init(d1_empty0 + i * 8, Cf::E1W);
      mbar_init(h_full0 + i * 8, 32 * Cf::E1W);
      mbar_init(h_empty0 + i * 8, 1);
      mbar_init(d2_full0 + i * 8, 1);
      mbar_init(d2_empty0 + i * 8, Cf::E2W);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)Cf::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_grid_sync();   // everything above is input-independent and overlaps the tail of the previous kernel
  for (int c = threadIdx.x; c < LDK; c += Cf::THREADS) {
    bs0[c] = b0 ? b0[c] : 0.f;
    bs3[c] = b3 ? b3[c] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(w_full, 2 * Cf::W_BYTES);
      for (int kb = 0; kb < NKB; ++kb) {
        tma_load_2d(smem_base + Cf::W0_OFF + kb * Cf::KB_W, &tmW0, w_full, kb * BK, 0);
        tma_load_2d(smem_base + Cf::W3_OFF + kb * Cf::KB_W, &tmW3, w_full, kb * BK, 0);
      }
      for (int i = 0; i < n_my; ++i) {
        const uint32_t sa = i % Cf::NA, ua = i / Cf::NA;
        const int panel = blockIdx.x + i * gridDim.x;
        mbar_wait(a_empty0 + sa * 8, (ua & 1) ^ 1);
        mbar_expect_tx(a_full0 + sa * 8, Cf::P_BYTES);
        for (int kb = 0; kb < NKB; ++kb)
          tma_load_2d(smem_base + Cf::A_OFF + sa * Cf::P_BYTES + kb * Cf::KB_P, &tmA, a_full0 + sa * 8, kb * BK, panel * BM);
      }
    }
  } else if (warp == 1) {
    // the whole warp runs this loop (warp-uniform control flow keeps descriptors in uniform registers); one elected
    // lane issues the MMAs and commits
    {
      constexpr uint32_t idesc = make_idesc(BM, LDK);
      mbar_wait(w_full, 0);
      auto mma1 = [&](int i) {
        const uint32_t st = i & 1, u = i >> 1, sa = i % Cf::NA, ua = i / Cf::NA;
        mbar_wait(a_full0 + sa * 8, ua & 1);
        mbar_wait(d1_empty0 + st * 8, (u & 1) ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + st * 2 * LDK;
        if (elect_one()) {
#pragma unroll
          for (int kb = 0; kb < NKB; ++kb) {
            const uint64_t adesc = make_sdesc(smem_base + Cf::A_OFF + sa * Cf::P_BYTES + kb * Cf::KB_P);
            const uint64_t bdesc = make_sdesc(smem_base + Cf::W0_OFF + kb * Cf::KB_W);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma_f16(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(d1_full0 + st * 8);
        }
        __syncwarp();
      };
      auto mma2 = [&](int i) {
        const uint32_t st = i & 1, u = i >> 1;
        mbar_wait(h_full0 + st * 8, u & 1);
        mbar_wait(d2_empty0 + st * 8, (u & 1) ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + st * 2 * LDK + LDK;
        if (elect_one()) {
#pragma unroll
          for (int kb = 0; kb < NKB; ++kb) {
            const uint64_t adesc = make_sdesc(smem_base + Cf::H_OFF + st * Cf::P_BYTES + kb * Cf::KB_P);
            const uint64_t bdesc = make_sdesc(smem_base + Cf::W3_OFF + kb * Cf::KB_W);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma_f16(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(d2_full0 + st * 8);
          umma_commit(h_empty0 + st * 8);
        }
        __syncwarp();
      };
      if (n_my > 0) mma1(0);
      for (int i = 0; i < n_my; ++i) {
        if (i + 1 < n_my) mma1(i + 1);
        mma2(i);
      }
    }
  } else if (warp < 2 + Cf::E1W) {
    // ---- epilogue 1: hidden = gelu(D1 + b0) -> bf16 H tile (K-major, SWIZZLE_128B) ----
    // lane quadrant = warp % 4, column half = (warp - 2) / 4
    const int lg = warp & 3, row = lg * 32 + lane;
    constexpr int NC1 = (LDK / 32) / (Cf::E1W / 4);
    const int c1_0 = ((warp - 2) >> 2) * NC1;
    for (int i = 0; i < n_my; ++i) {
      const uint32_t st = i & 1, u = i >> 1;
      mbar_wait(d1_full0 + st * 8, u & 1);
      mbar_wait(h_empty0 + st * 8, (u & 1) ^ 1);
      tc_fence_after();
      uint8_t* hrow = smem + Cf::H_OFF + st * Cf::P_BYTES + row * 128;
#pragma unroll 1
      for (int c4 = c1_0; c4 < c1_0 + NC1; ++c4) {
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(st * 2 * LDK + c4 * 32), r);
        if (c4 == c1_0 + NC1 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(d1_empty0 + st * 8);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 pk;
          uint32_t* pw = reinterpret_cast<uint32_t*>(&pk);
          // bias of 8 columns as two broadcast LDS.128 (a 32-bit load per element would cost a wavefront each)
          const float4 ba = *reinterpret_cast<const float4*>(bs0 + c4 * 32 + j * 8), bb = *reinterpret_cast<const float4*>(bs0 + c4 * 32 + j * 8 + 4);
          const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const __nv_bfloat162 t = __floats2bfloat162_rn(gelu_tanh(__uint_as_float(r[j * 8 + 2 * e]) + bv[2 * e]),
                                                          gelu_tanh(__uint_as_float(r[j * 8 + 2 * e + 1]) + bv[2 * e + 1]));
            pw[e] = *reinterpret_cast<const uint32_t*>(&t);
          }
          const int c = c4 * 4 + j;                               // 16-byte chunk of the row
          *reinterpret_cast<uint4*>(hrow + (c >> 3) * Cf::KB_P + (((c & 7) ^ (row & 7)) << 4)) = pk;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(h_full0 + st * 8);
    }
  } else {
    // ---- epilogue 2: out = k + D2 + b3.  k is read back from the TMA tile and the result overwrites it in place (same
    //      swizzled position, every thread owns its row); the finished tile leaves with one TMA store per k-block, so
    //      the global writes are full lines instead of 32 row-strided 16-byte pieces per instruction ----
    const int lg = warp & 3, row = lg * 32 + lane;
    constexpr int NC2 = (LDK / 32) / (Cf::E2W / 4);
    const int c2_0 = ((warp - 2 - Cf::E1W) >> 2) * NC2;
    for (int i = 0; i < n_my; ++i) {
      const uint32_t st = i & 1, u = i >> 1;
      const int panel = blockIdx.x + i * gridDim.x;
      mbar_wait(d2_full0 + st * 8, u & 1);
      tc_fence_after();
      const uint32_t sa = i % Cf::NA;
      uint8_t* arow = smem + Cf::A_OFF + sa * Cf::P_BYTES + row * 128;
#pragma unroll 1
      for (int c4 = c2_0; c4 < c2_0 + NC2; ++c4) {
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(st * 2 * LDK + LDK + c4 * 32), r);
        if (c4 == c2_0 + NC2 - 1) {                                // this warp's part of the accumulator has been read
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(d2_empty0 + st * 8);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = c4 * 4 + j;
          uint4* slot = reinterpret_cast<uint4*>(arow + (c >> 3) * Cf::KB_P + (((c & 7) ^ (row & 7)) << 4));
          const uint4 res = *slot;
          const __nv_bfloat162* kh = reinterpret_cast<const __nv_bfloat162*>(&res);
          uint4 pk;
          uint32_t* pw = reinterpret_cast<uint32_t*>(&pk);
          const float4 ba = *reinterpret_cast<const float4*>(bs3 + c4 * 32 + j * 8), bb = *reinterpret_cast<const float4*>(bs3 + c4 * 32 + j * 8 + 4);
          const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 kv = __bfloat1622float2(kh[e]);
            const __nv_bfloat162 t = __floats2bfloat162_rn(kv.x + __uint_as_float(r[j * 8 + 2 * e]) + bv[2 * e],
                                                          kv.y + __uint_as_float(r[j * 8 + 2 * e + 1]) + bv[2 * e + 1]);
            pw[e] = *reinterpret_cast<const uint32_t*>(&t);
          }
          *slot = pk;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic writes -> visible to the TMA store
      asm volatile("bar.sync 3, %0;" ::"n"(32 * Cf::E2W) : "memory");   // the epilogue-2 warps
      if (warp == 2 + Cf::E1W && lane == 0) {
        for (int kb = 0; kb < NKB; ++kb)
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                       ::"l"(&tmOut), "r"(smem_base + Cf::A_OFF + sa * Cf::P_BYTES + kb * Cf::KB_P), "r"(kb * BK), "r"(panel * BM)
                       : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the slot has been read: the producer may refill it
        mbar_arrive(a_empty0 + sa * 8);
      }
    }
    if (warp == 2 + Cf::E1W && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cf::TMEM_COLS) : "memory");
  }
}

template <int LDK>
int launch_kernel_fixup(const void* k, int lda, const void* W0, int ldw0, const float* b0, const void* W3, int ldw3,
                        const float* b3, int M, void* out, int ldo, cudaStream_t st) {
  using Cf = FxCfg<LDK>;
  CUtensorMap ta, tw0, tw3;
  int rc = make_map(&ta, k, M, LDK, lda, BM);
  if (rc) return rc;
  rc = make_map(&tw0, W0, LDK, LDK, ldw0, LDK);
  if (rc) return rc;
  rc = make_map(&tw3, W3, LDK, LDK, ldw3, LDK);
  if (rc) return rc;
  CUtensorMap tout;
  rc = make_map(&tout, out, M, LDK, ldo, BM);
  if (rc) return rc;
  CSEG_SET_SMEM(kernel_fixup_kernel<LDK>, Cf::SMEM_BYTES);
  const int panels = cdiv(M, BM);
  cseg_launch(kernel_fixup_kernel<LDK>, dim3(std::min(panels, sm_count())), dim3(Cf::THREADS), Cf::SMEM_BYTES, st, ta, tw0,
              tw3, tout, M, b0, b3);
  CSEG_LAUNCH_CHECK("jbu_kernel_fixup");
  return 0;
}

}  // namespace

// returns 1 when the shape is not covered by the fused kernel (caller runs cseg_gemm + cseg_norm_sim instead)
int cseg_fixup_norm_sim_tc(const void* y, int ldy, const void* W, int ldw, int M, int C, const float* bias, float alpha,
                           const float* text, int Q, const float* cls_bias, int hw, float* logits, cudaStream_t st) {
  if (C % 128 != 0 || C > 512 || Q > 16 || ldy % 8 != 0 || ldw % 8 != 0) return 1;
  CUtensorMap ta, tb;
  int rc = make_map(&ta, y, M, C, ldy, BM);
  if (rc) return rc;
  rc = make_map(&tb, W, C, C, ldw, 128);
  if (rc) return rc;
  if (Q <= 8)
    return launch_normsim<8>(ta, tb, M, C, (const bf16*)y, ldy, bias, alpha, text, Q, cls_bias, hw, logits, st);
  return launch_normsim<16>(ta, tb, M, C, (const bf16*)y, ldy, bias, alpha, text, Q, cls_bias, hw, logits, st);
}

int cseg_basis_logits_tc(const void* s, int lds, int Cb, int n_crops, int hw, int T, int tstride, const void* gram,
                         const void* aux, int ldg, const float* consts, int Q, const float* cls_bias, float* logits,
                         cudaStream_t st) {
  CSEG_REQUIRE(n_crops > 0 && hw > 0 && hw % BM == 0, "basis_logits: hw=%d must be a positive multiple of %d", hw, BM);
  CSEG_REQUIRE(Q > 0 && Q <= 31, "basis_logits: Q=%d must be in 1..31", Q);
  const int n2 = Q + 1 <= 16 ? 16 : 32;                 // aux rows: Q text rows + the bias row
  CSEG_REQUIRE(T > 0 && T <= 256 - n2 && tstride >= T && tstride % 8 == 0,
               "basis_logits: T=%d (tokens per crop) must be in 1..%d, tstride=%d >= T and a multiple of 8 (TMA)", T, 256 - n2, tstride);
  const int KS = cdiv(T, 16);
  CSEG_REQUIRE(KS * 16 + n2 <= 256, "basis_logits: T=%d with Q=%d needs %d accumulator columns (max 256)", T, Q, KS * 16 + n2);
  CSEG_REQUIRE(Cb >= KS * 16 && lds >= Cb, "basis_logits: coefficient rows need >= %d columns (Cb=%d, lds=%d)", KS * 16, Cb, lds);
  CSEG_REQUIRE(lds % 8 == 0 && ldg % 8 == 0, "basis_logits: lds=%d, ldg=%d must be multiples of 8", lds, ldg);
  CSEG_REQUIRE(((uintptr_t)s & 15) == 0 && ((uintptr_t)gram & 15) == 0 && ((uintptr_t)aux & 15) == 0,
               "basis_logits: operands must be 16-byte aligned");
  const long long M = (long long)n_crops * hw;
  CSEG_REQUIRE(M < (1ll << 31), "basis_logits: too many pixels");
  CUtensorMap ta, tb1, tb2;
  int rc = make_map(&ta, s, (int)M, Cb, lds, BM);
  if (rc) return rc;
  rc = make_map(&tb1, gram, n_crops * tstride, ldg, ldg, KS * 16);
  if (rc) return rc;
  rc = make_map(&tb2, aux, n2, ldg, ldg, n2);
  if (rc) return rc;
  CSEG_SET_SMEM(basis_logits_kernel, BlCfg::SMEM_BYTES);
  const int panels = (int)(M / BM);
  cseg_launch(basis_logits_kernel, dim3(std::min(panels, sm_count())), dim3(BlCfg::THREADS), BlCfg::SMEM_BYTES, st, 
      ta, tb1, tb2, panels, hw, tstride, KS, (const bf16*)s, lds, consts, Q, cls_bias, logits, n2);
  CSEG_LAUNCH_CHECK("basis_logits");
  return 0;
}

// returns 1 when the shape is not covered (caller issues the two GEMMs instead)
int cseg_jbu_kernel_fixup_tc(const void* k, int lda, const void* W0, int ldw0, const float* b0, const void* W3, int ldw3,
                             const float* b3, int M, int ldk, void* out, int ldo, cudaStream_t st) {
  if ((ldk != 64 && ldk != 128) || lda % 8 || ldw0 % 8 || ldw3 % 8 || ldo % 8) return 1;
  if ((((uintptr_t)k | (uintptr_t)W0 | (uintptr_t)W3 | (uintptr_t)out) & 15) != 0) return 1;
  if (ldk == 128) return launch_kernel_fixup<128>(k, lda, W0, ldw0, b0, W3, ldw3, b3, M, out, ldo, st);
  return launch_kernel_fixup<64>(k, lda, W0, ldw0, b0, W3, ldw3, b3, M, out, ldo, st);
}

template <int ACT, int OUTB, int RES>
int launch2cta3(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const EpiParams& ep, cudaStream_t st) {
  constexpr int BN2 = 256, ST2 = 4;
  using C = Cfg2<BN2, ST2>;
  CSEG_SET_SMEM((gemm_bf16_tcgen05_2cta_kernel<BN2, ST2, ACT, OUTB, RES>), C::SMEM_BYTES);
  const int m2_tiles = cdiv(M, 2 * BM), num_tiles = m2_tiles * (N / BN2);
  const int grid = 2 * std::min(num_tiles, sm_count() / 2);
  cseg_launch(gemm_bf16_tcgen05_2cta_kernel<BN2, ST2, ACT, OUTB, RES>, dim3(grid), dim3(C::THREADS), C::SMEM_BYTES, st, ta, tb, K,
              m2_tiles, num_tiles, ep);
  CSEG_LAUNCH_CHECK("gemm_bf16_tcgen05_2cta");
  return 0;
}
template <int ACT>
int launch2cta1(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const EpiParams& ep, cudaStream_t st) {
  if (ep.out_bf16) return ep.residual ? launch2cta3<ACT, 1, 1>(ta, tb, M, N, K, ep, st) : launch2cta3<ACT, 1, 0>(ta, tb, M, N, K, ep, st);
  return ep.residual ? launch2cta3<ACT, 0, 1>(ta, tb, M, N, K, ep, st) : launch2cta3<ACT, 0, 0>(ta, tb, M, N, K, ep, st);
}
bool gemm_2cta_enabled() {     // CSEG_GEMM_2CTA=0 selects the single-CTA kernel for every shape (A/B measurements)
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("CSEG_GEMM_2CTA");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

// Per-block Gram matrices of fp32-grade accuracy on the bf16 tensor cores: X is the [hi | lo] bf16 split of a row-normalised
// fp32 matrix ([M, 2 * width], width % 64 == 0); value(b, i, j) = alpha * <x_(b,i), x_(b,j)> for the square diagonal blocks
// of block_rows rows and i, j >= skip.
//   layout 0: compact fp32 [n_blocks, block_rows - skip, block_rows - skip] (the similarity map layout)
//   layout 1: fp32 [n_blocks, ceil(cols_pad / 32), cols_pad, 32] indexed [b][i / 32][j][i % 32] with UNSHIFTED i, j (entries with
//             i < skip or j < skip are never written): what a TMEM-row-per-lane consumer reads with coalesced loads
int cseg_gram_split_tc(const void* X, int ldx, int M, int width, int block_rows, int skip, float alpha, float* out,
                       int layout, int cols_pad, cudaStream_t st) {
  CSEG_REQUIRE(M > 0 && width > 0 && width % BK == 0 && ldx % 8 == 0 && ldx >= 2 * width, "gram_split: M=%d width=%d ldx=%d", M,
               width, ldx);
  CSEG_REQUIRE(block_rows > skip && skip >= 0 && M % block_rows == 0, "gram_split: M=%d block_rows=%d skip=%d", M, block_rows, skip);
  CSEG_REQUIRE(((uintptr_t)X & 15) == 0, "gram_split: operand must be 16-byte aligned");
  CSEG_REQUIRE(layout == 0 || (layout == 1 && cols_pad >= block_rows), "gram_split: layout=%d cols_pad=%d", layout, cols_pad);
  CUtensorMap ta, tb;
  int rc = make_map(&ta, X, M, 2 * width, ldx, BM);
  if (rc) return rc;
  rc = make_map(&tb, X, M, 2 * width, ldx, 128);
  if (rc) return rc;
  const int P = block_rows - skip;
  EpiParams ep{nullptr, nullptr, 0, 0, alpha, CSEG_ACT_NONE, 0, out, 0, M, M, block_rows, (long long)P * P, P, skip, 0, width / BK};
  if (layout == 1) {
    ep.cmp_mode = 1;
    ep.cmp_rs = cols_pad;
    ep.cmp_cs = (long long)((cols_pad + 31) / 32) * cols_pad * 32;
  }
  return launch3<128, 4, CSEG_ACT_NONE, 0, 3>(ta, tb, M, M, 3 * width, ep, st);
}

int cseg_gemm_bf16_tc(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const float* bias,
                      const void* residual, int ldr, int res_dtype, float alpha, int act, int out_dtype, void* C,
                      int ldc, cudaStream_t st, int diag_rows) {
  CSEG_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  CSEG_REQUIRE(K % 8 == 0, "gemm(bf16): K=%d must be a multiple of 8 (16-byte rows for TMA)", K);
  CSEG_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "gemm(bf16): lda=%d, ldb=%d must be multiples of 8", lda, ldb);
  CSEG_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0, "gemm(bf16): operands must be 16-byte aligned");
  // Tile width: the widest BN whose tile count still gives every SM work; ties broken by the makespan
  // (rounds x tile width).  BN = 192 / 256 halve the operand traffic per flop of the large-N GEMMs.
  const int m_tiles = cdiv(M, BM), sms = sm_count();
  int bn = 64;
  {
    double best = 1e30;
    const int cands[4] = {256, 192, 128, 64};
    for (int ci = 0; ci < 4; ++ci) {
      const int c = cands[ci];
      if (c > 64 && N < c) continue;
      const long long tiles = (long long)m_tiles * cdiv(N, c);
      const long long rounds = (tiles + sms - 1) / sms;
      // cost model: rounds x (tile MMA time + feed penalty for narrow tiles)
      const double feed = (c == 64) ? 1.6 : (c == 128 ? 1.25 : 1.0);
      // BN = 192 has 12 epilogue warps for 6 chunks: activation epilogues (MUFU / issue bound) pace its tiles
      const double epi = (c == 192 && act != CSEG_ACT_NONE) ? 1.45 : 1.0;
      const double cost = (double)rounds * c * feed * epi;
      if (cost < best - 1e-9) { best = cost; bn = c; }
    }
  }
  CUtensorMap ta, tb;
  int rc = make_map(&ta, A, M, K, lda, BM);
  if (rc) return rc;
  EpiParams ep{bias, residual, ldr, res_dtype == CSEG_BF16, alpha, act, out_dtype == CSEG_BF16, C, ldc, M, N, diag_rows, 0, 0, 0, 0, 0};
  // CTA pairs (cta_group::2) for the big N = 256 k GEMMs of the ViT: the pair stages 32 KB per k-block and SM instead of 48
  // (measured at M = 18 912: QKV 59.5 -> 53.6 us, fc2 82.4 -> 74.6 us; the K = 768 residual GEMM (out-proj) is paced by its
  // epilogue and loses 8 % to the pair's lock step, the GELU GEMM (fc1) is epilogue bound either way: both stay single-CTA)
  const bool pair_shape = residual == nullptr || (act == CSEG_ACT_NONE && K >= 1536);
  if (bn == 256 && N % 256 == 0 && diag_rows == 0 && (residual == nullptr || res_dtype == CSEG_F32) && sms % 2 == 0 && pair_shape &&
      (long long)cdiv(M, 2 * BM) * (N / 256) >= sms / 2 && gemm_2cta_enabled()) {
    rc = make_map(&tb, B, N, K, ldb, 128);
    if (rc) return rc;
    if (act == CSEG_ACT_GELU) return launch2cta1<CSEG_ACT_GELU>(ta, tb, M, N, K, ep, st);
    if (act == CSEG_ACT_QUICKGELU) return launch2cta1<CSEG_ACT_QUICKGELU>(ta, tb, M, N, K, ep, st);
    return launch2cta1<CSEG_ACT_NONE>(ta, tb, M, N, K, ep, st);
  }
  rc = make_map(&tb, B, N, K, ldb, bn);
  if (rc) return rc;
  if (bn == 64) return launch<64, 6>(ta, tb, M, N, K, ep, st);
  if (bn == 192) return launch<192, 4>(ta, tb, M, N, K, ep, st);
  if (bn == 256) return launch<256, 3>(ta, tb, M, N, K, ep, st);
  return launch<128, 4>(ta, tb, M, N, K, ep, st);
}
