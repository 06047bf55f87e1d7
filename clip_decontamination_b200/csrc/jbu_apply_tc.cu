// JBU adaptive convolution on the 5th-gen tensor cores (bf16 path of cseg_jbu_apply, C % 128 == 0).
//
//   out[y, x, c] = sum_{i,j} hr[reflect(y+i-R), reflect(x+j-R), c] * kern[(y,x)][i*D + j]        D = 2R+1
//
// (simfeatup_dev/upsamplers.py:269-274, semantics of adaptive_conv_py_simple :14-25.)
//
// GEMM view, per source row s of a tile of RW = 4 output rows x 16 pixels and a slab of 128 channels:
//   D[128 ch, 64 px] += A_s[128 ch, 32 pos] . B_s[32 pos, 64 px]
//   A_s = hr[row s, positions x0-8 .. x0+23, channels]      (the tile's source strip, [pos][ch] as stored in HBM)
//   B_s[k][(r, m)] = kern[(y0+r, x0+m)][(s-r)*D + j]  with k = m + j + 8 - R  (zero where s-r or j is out of range)
// i.e. the kernel weights are expanded into banded matrices and every source row costs two tcgen05.mma
// (M=128, N=64, K=16) per channel slab.  The channel dimension is M, so the feature map is consumed exactly
// in its HBM layout (MN-major A operand, SWIZZLE_128B) and the accumulator holds 64 pixels x 128 channels in TMEM.
//
// Warp roles (persistent CTA, one per SM, tiles x-fastest):
//   warps 0-3   epilogue: tcgen05.ld (lane = channel) -> bf16 -> dst[pixel][channel]
//   warp  4     TMEM allocator + MMA issuer
//   warps 5-8   A loaders: cp.async 16-byte chunks (reflect padding resolved per chunk) into a 4-stage ring of
//               source strips, written directly in the swizzled MN-major layout
//   warps 9-12  B builders: load the tile's 64 x D*D weights, scatter them into the banded K-major tiles
//               (double-buffered; the zero background is written once per buffer, the band positions do not
//               depend on the data)
// All hand-offs are mbarriers; generic-proxy writes are fenced (fence.proxy.async) before the tensor core reads.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int CV_RW = 4, CV_TX = 16, CV_NPX = CV_RW * CV_TX;   // 64 pixels per tile = N of the MMA
constexpr int CV_NPOS = 32;                                    // source positions per strip (K per source row)
constexpr int CV_NSTG = 6;                                     // A ring stages
constexpr int CV_INFL = 4;                                     // source strips in flight per loader thread
constexpr int CV_EPI_WARPS = 4, CV_LD_WARPS = 4, CV_BB_WARPS = 4;
constexpr int CV_THREADS = 32 * (CV_EPI_WARPS + 1 + CV_LD_WARPS + CV_BB_WARPS);
constexpr int CV_LD_T0 = 32 * (CV_EPI_WARPS + 1), CV_BB_T0 = CV_LD_T0 + 32 * CV_LD_WARPS;

__device__ __forceinline__ int cv_reflect(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// MN-major SWIZZLE_128B descriptor (cute make_umma_desc<Major::MN>, LayoutType::B128):
//   ((8 elem, 8, m), (8, k)) : ((1, 8, LBO), (64 elem = 128 B, SBO))
// 64 contiguous M elements per 128 B row, K rows 128 B apart, 8-row groups SBO apart, 64-element M chunks LBO apart.
__device__ __forceinline__ uint64_t cv_adesc(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: D f32, A/B bf16, A MN-major (bit 15), B K-major, N = 64, M = 128
__host__ __device__ constexpr uint32_t cv_idesc() {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((uint32_t)(CV_NPX >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

template <int R, int MH>
struct CvCfg {
  static constexpr int D = 2 * R + 1, NSRC = CV_RW + 2 * R, NPAIR = (NSRC + 1) / 2;
  static constexpr int CH = 128 * MH;                          // channels per tile
  static constexpr int CHUNK_BYTES = CV_NPOS * 128;            // one 64-channel chunk of a strip (32 rows x 128 B)
  static constexpr int A_STAGE = (CH / 64) * CHUNK_BYTES;      // 8 / 16 KB
  static constexpr int B_PAIR = CV_NPX * 128;                  // [64 px][64 k] bf16: two source rows per 128 B row
  static constexpr int B_BUF = NPAIR * B_PAIR;
  static constexpr int A_OFF = 0, B_OFF = CV_NSTG * A_STAGE, BAR_OFF = B_OFF + 2 * B_BUF;
  static constexpr int NBARS = 2 * CV_NSTG + 8;
  static constexpr int SMEM_BYTES = BAR_OFF + NBARS * 8 + 16 + 1024;
  static constexpr int TMEM_COLS = (2 * MH * CV_NPX <= 128) ? 128 : 256;
};

template <int R, int MH>
__global__ void __launch_bounds__(CV_THREADS, 1)
adaptive_conv_tc_kernel(const bf16* __restrict__ hr, int H2, int W2, int C, const bf16* __restrict__ kern, int ldk,
                        bf16* __restrict__ dst, int nx, int ny, int nslab, int total_tiles) {
  using Cf = CvCfg<R, MH>;
  constexpr int D = Cf::D, NSRC = Cf::NSRC;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic on the __shared__ array: accesses compile to LDS / STS, not generic LD / ST
  uint64_t* bars = (uint64_t*)(smem + Cf::BAR_OFF);
  uint32_t* tmem_slot = (uint32_t*)(bars + Cf::NBARS);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t a_full0 = smem_u32(bars), a_empty0 = a_full0 + CV_NSTG * 8;
  const uint32_t b_full0 = a_empty0 + CV_NSTG * 8, b_empty0 = b_full0 + 16;
  const uint32_t t_full0 = b_empty0 + 16, t_empty0 = t_full0 + 16;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < CV_NSTG; ++i) {
      mbar_init(a_full0 + i * 8, 32 * CV_LD_WARPS);
      mbar_init(a_empty0 + i * 8, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(b_full0 + i * 8, 32 * CV_BB_WARPS);
      mbar_init(b_empty0 + i * 8, 1);
      mbar_init(t_full0 + i * 8, 1);
      mbar_init(t_empty0 + i * 8, CV_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == CV_EPI_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)Cf::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_grid_sync();   // everything above is input-independent and overlaps the tail of the previous kernel
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile id -> coordinates (x fastest, then y, then crop x channel slab)
  auto tile_coords = [&](int tile, int& x0, int& y0, int& crop, int& c0) {
    const int xt = tile % nx, rest = tile / nx;
    const int yt = rest % ny, z = rest / ny;
    x0 = xt * CV_TX;
    y0 = yt * CV_RW;
    crop = z / nslab;
    c0 = (z % nslab) * Cf::CH;
  };

  if (warp < CV_EPI_WARPS) {
    // ---------------- epilogue: lane = channel, 64 pixel columns per channel slab ----------------
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
      int x0, y0, crop, c0;
      tile_coords(tile, x0, y0, crop, c0);
      const uint32_t as = tl & 1, aph = (tl >> 1) & 1;
      mbar_wait(t_full0 + as * 8, aph);
      tc_fence_after();
#pragma unroll 1
      for (int hc = 0; hc < MH * (CV_NPX / 32); ++hc) {          // 32 pixel columns (two output rows) per step
        const int half = hc / (CV_NPX / 32), cb = hc % (CV_NPX / 32);
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)((as * MH + half) * CV_NPX + cb * 32), r);
        bf16* obase = dst + (size_t)crop * H2 * W2 * C + c0 + half * 128 + warp * 32 + lane;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int y = y0 + cb * 2 + rr;
          if (y >= H2) continue;
          bf16* o = obase + ((size_t)y * W2 + x0) * C;
          const int mmax = min(16, W2 - x0);
          if (mmax == 16) {
#pragma unroll
            for (int m = 0; m < 16; ++m) o[(size_t)m * C] = __float2bfloat16_rn(__uint_as_float(r[rr * 16 + m]));
          } else {
#pragma unroll
            for (int m = 0; m < 16; ++m)
              if (m < mmax) o[(size_t)m * C] = __float2bfloat16_rn(__uint_as_float(r[rr * 16 + m]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_empty0 + as * 8);
    }
  } else if (warp == CV_EPI_WARPS) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      constexpr uint32_t idesc = cv_idesc();
      uint32_t tl = 0, it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
        const uint32_t as = tl & 1, aph = (tl >> 1) & 1;
        mbar_wait(t_empty0 + as * 8, aph ^ 1);
        mbar_wait(b_full0 + as * 8, aph);
        tc_fence_after();
        const uint32_t bbuf = smem_base + Cf::B_OFF + as * Cf::B_BUF;
        for (int s = 0; s < NSRC; ++s, ++it) {
          const uint32_t st = it % CV_NSTG, ph = (it / CV_NSTG) & 1;
          mbar_wait(a_full0 + st * 8, ph);
          tc_fence_after();
          const uint32_t a_src = smem_base + Cf::A_OFF + st * Cf::A_STAGE;
#pragma unroll
          for (int half = 0; half < MH; ++half) {
            const uint32_t tacc = tmem_base + (uint32_t)((as * MH + half) * CV_NPX);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint64_t adesc = cv_adesc(a_src + half * 2 * Cf::CHUNK_BYTES + ks * 16 * 128, Cf::CHUNK_BYTES);
              const uint64_t bdesc = make_sdesc(bbuf + (s >> 1) * Cf::B_PAIR + ((s & 1) * 2 + ks) * 32);
              umma_f16(tacc, adesc, bdesc, idesc, (s > 0 || ks > 0) ? 1u : 0u);
            }
          }
          umma_commit(a_empty0 + st * 8);
        }
        umma_commit(t_full0 + as * 8);
        umma_commit(b_empty0 + as * 8);
      }
    }
  } else if (warp < CV_EPI_WARPS + 1 + CV_LD_WARPS) {
    // ---------------- A loaders: source strips into the ring ----------------
    const int lt = tid - CV_LD_T0;                               // 0..127
    constexpr int CPR = Cf::CH / 8;                              // 16-byte chunks per position
    constexpr int LPT = CV_NPOS * CPR / (32 * CV_LD_WARPS);      // chunks per thread per strip (4 / 8)
    uint32_t it = 0;                                             // strips issued so far; strip j lives in stage j % NSTG
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int x0, y0, crop, c0;
      tile_coords(tile, x0, y0, crop, c0);
      const bf16* hrc = hr + (size_t)crop * H2 * W2 * C + c0;
      int src_off[LPT];
      uint32_t dst_off[LPT];
#pragma unroll
      for (int k = 0; k < LPT; ++k) {
        const int e = lt + k * 32 * CV_LD_WARPS;
        const int p = e / CPR, c8 = e % CPR;
        const int xx = cv_reflect(min(max(x0 - 8 + p, -(W2 - 1)), 2 * (W2 - 1)), W2);
        src_off[k] = xx * C + c8 * 8;
        dst_off[k] = (uint32_t)((c8 >> 3) * Cf::CHUNK_BYTES + p * 128 + (((c8 & 7) ^ (p & 7)) << 4));
      }
      for (int s = 0; s < NSRC; ++s, ++it) {
        const uint32_t st = it % CV_NSTG, ph = (it / CV_NSTG) & 1;
        mbar_wait(a_empty0 + st * 8, ph ^ 1);
        const int yy = cv_reflect(min(max(y0 + s - R, -(H2 - 1)), 2 * (H2 - 1)), H2);
        const bf16* rowp = hrc + (size_t)yy * W2 * C;
        const uint32_t base = smem_base + Cf::A_OFF + st * Cf::A_STAGE;
#pragma unroll
        for (int k = 0; k < LPT; ++k)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(base + dst_off[k]), "l"(rowp + src_off[k]) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (it >= CV_INFL - 1) {                                 // strip it-(INFL-1) has landed: publish it
          asm volatile("cp.async.wait_group %0;" ::"n"(CV_INFL - 1) : "memory");
          fence_proxy_async();
          mbar_arrive(a_full0 + ((it - (CV_INFL - 1)) % CV_NSTG) * 8);
        }
      }
    }
    // drain: the last INFL-1 strips, oldest first
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    fence_proxy_async();
    for (uint32_t j = (it >= CV_INFL - 1 ? it - (CV_INFL - 1) : 0); j < it; ++j) mbar_arrive(a_full0 + (j % CV_NSTG) * 8);
  } else {
    // ---------------- B builders: banded weight tiles ----------------
    const int bt = tid - CV_BB_T0;                               // 0..127
    constexpr int WPT = CV_NPX * 16 / (32 * CV_BB_WARPS);        // 8 x 16-byte weight chunks per thread
    // the 16-byte chunk index v = bt & 15 (hence the 8 taps a thread scatters) is the same for all of its chunks
    const int v = bt & 15;
    int tap_i[8], tap_j[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int t = v * 8 + q;
      tap_i[q] = (t < D * D) ? t / D : -1;
      tap_j[q] = t - (t / D) * D;
    }
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
      int x0, y0, crop, c0;
      tile_coords(tile, x0, y0, crop, c0);
      const uint32_t as = tl & 1, aph = (tl >> 1) & 1;
      uint4 wv[WPT];
#pragma unroll
      for (int k = 0; k < WPT; ++k) {
        const int e = bt + k * 32 * CV_BB_WARPS;
        const int n = e >> 4, y = y0 + (n >> 4), x = x0 + (n & 15);
        wv[k] = make_uint4(0, 0, 0, 0);
        if (y < H2 && x < W2 && v * 8 < ldk)
          wv[k] = __ldg(reinterpret_cast<const uint4*>(kern + (((size_t)crop * H2 + y) * W2 + x) * ldk + v * 8));
      }
      mbar_wait(b_empty0 + as * 8, aph ^ 1);
      uint8_t* bbuf = smem + Cf::B_OFF + as * Cf::B_BUF;
      if (tl < 2) {                                              // zero background, once per buffer
        for (int e = bt; e < Cf::B_BUF / 16; e += 32 * CV_BB_WARPS) reinterpret_cast<uint4*>(bbuf)[e] = make_uint4(0, 0, 0, 0);
        asm volatile("bar.sync 2, %0;" ::"n"(32 * CV_BB_WARPS) : "memory");
      }
#pragma unroll
      for (int k = 0; k < WPT; ++k) {
        const int e = bt + k * 32 * CV_BB_WARPS;
        const int n = e >> 4, r = n >> 4, m = n & 15;
        const unsigned short* hv = reinterpret_cast<const unsigned short*>(&wv[k]);
        uint8_t* rowb = bbuf + (n >> 3) * 1024 + (n & 7) * 128;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (tap_i[q] < 0) continue;
          const int s = r + tap_i[q];                            // source row fed by this tap
          const int kk = (s & 1) * 32 + m + tap_j[q] + 8 - R;    // column inside the 64-wide (two source rows) tile row
          *reinterpret_cast<unsigned short*>(rowb + (s >> 1) * Cf::B_PAIR + ((((kk >> 3) ^ (n & 7)) << 4) | ((kk & 7) << 1))) = hv[q];
        }
      }
      fence_proxy_async();
      mbar_arrive(b_full0 + as * 8);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == CV_EPI_WARPS) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cf::TMEM_COLS) : "memory");
  }
}

template <int R, int MH>
int launch_conv_tc(const bf16* hr, int n_crops, int H2, int W2, int C, const bf16* kern, int ldk, bf16* dst, cudaStream_t st) {
  using Cf = CvCfg<R, MH>;
  CSEG_SET_SMEM((adaptive_conv_tc_kernel<R, MH>), Cf::SMEM_BYTES);
  const int nx = cdiv(W2, CV_TX), ny = cdiv(H2, CV_RW), nslab = C / Cf::CH;
  const long long total = (long long)nx * ny * n_crops * nslab;
  CSEG_REQUIRE(total < (1ll << 31), "jbu_apply(bf16): too many tiles");
  const int grid = (int)std::min<long long>(total, sm_count());
  cseg_launch(adaptive_conv_tc_kernel<R, MH>, dim3(grid), dim3(CV_THREADS), Cf::SMEM_BYTES, st, hr, H2, W2, C, kern, ldk,
              dst, nx, ny, nslab, (int)total);
  CSEG_LAUNCH_CHECK("jbu_adaptive_conv_tc");
  return 0;
}

}  // namespace

// returns 1 when the shape is not covered (caller falls back to the mma.sync kernel)
int cseg_jbu_adaptive_conv_tc(const bf16* hr, int n_crops, int H2, int W2, int C, const bf16* kern, int ldk, int radius,
                              bf16* dst, cudaStream_t st) {
  if (C % 128 != 0 || ldk % 8 != 0 || (radius != 5 && radius != 3)) return 1;
  if (H2 < 2 * radius + 2 || W2 < 24) return 1;                 // single reflection only
  if (((uintptr_t)hr & 15) != 0 || ((uintptr_t)kern & 15) != 0) return 1;
  if (C % 256 == 0) {
    if (radius == 5) return launch_conv_tc<5, 2>(hr, n_crops, H2, W2, C, kern, ldk, dst, st);
    return launch_conv_tc<3, 2>(hr, n_crops, H2, W2, C, kern, ldk, dst, st);
  }
  if (radius == 5) return launch_conv_tc<5, 1>(hr, n_crops, H2, W2, C, kern, ldk, dst, st);
  return launch_conv_tc<3, 1>(hr, n_crops, H2, W2, C, kern, ldk, dst, st);
}
