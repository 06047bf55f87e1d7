// JBU apply = bicubic x2 upsample + reflect pad + adaptive convolution, as ONE tcgen05 kernel on the LOW-resolution
// source (bf16 path of cseg_jbu_apply, C % 128 == 0).
//
//   out[y, x, c] = sum_{i,j} k[(y,x)][i*D + j] * hr[reflect(y+i-R), reflect(x+j-R), c]        D = 2R+1
//   hr[qy, qx, c] = sum_{a,b} cy_a(qy) cx_b(qx) src[clamp(..), clamp(..), c]                  (bicubic, A = -0.75)
//
// (simfeatup_dev/upsamplers.py:268-274; torch upsample_bicubic2d, align_corners=False.)  Both steps are linear and
// separable in the spatial position, so for every output pixel the D x D kernel over the high-res neighbourhood folds
// into a composite kernel over a (2 DO + 1)^2 LOW-res neighbourhood (DO = 4 for R = 5):
//   K'[(y,x)][dy, dx] = sum_{i,j} TY[y][i][dy] * k[(y,x)][i, j] * TX[x][j][dx]
// where TX[x][j][dx] is the bicubic weight that tap j of column x puts on low-res column (x >> 1) + dx - DO (reflect
// padding and border clamps are resolved when the table is built; TY likewise).  The high-res intermediate never
// exists: HBM traffic drops from (write + re-read hr, 5 x the source) to the source itself, and the banded GEMM sees
// 16 instead of 32 positions per source row and NLR = 10 instead of 14 rows per tile.
//
// Work unit = a vertical run of `seg` tiles of one (crop, 16-px column, channel slab): consecutive tiles of a run share
// NLR - 2 of their NLR low-res source rows, so the source strips live in a ring of 16 slots and only the 2 new rows per
// tile are fetched (the strips were 62 % of the kernel's L2 traffic: 10x halo amplification per 4 x 16 tile).
// Per tile of 8 rows x 16 px and 128-channel slab, every low-res row s of the (NLR x 16) source patch costs one
// tcgen05.mma  D[128 ch, 128 px] += A_s[128 ch, 16 pos] . B_s[16 pos, 128 px]  (A: MN-major SWIZZLE_128B straight from
// the [pos][ch] HBM layout; B: composite band tile, K-major).  Warp roles of the persistent CTA:
//   warps 0-7   epilogue (tcgen05.ld, lane = channel; lane quadrant = warp % 4, column half = warp / 4)
//   warp 8      TMEM allocator + MMA issuer
//   warps 9-10  strip loaders (cp.async into a 16-slot ring)
//   warps 11-18 band builders: two threads per pixel scatter its (2 DO + 1)^2 composite weights (computed by the
//               fz_composite pre-pass, fetched one tile ahead) into the double-buffered band tiles; the swizzled offsets
//               depend on (pixel, tap) only and are computed once per kernel.
#include "common.cuh"
#include "tc_common.cuh"
#include "jbu_share.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>

namespace {

constexpr int FZ_RW = 8, FZ_TX = 16, FZ_NPX = FZ_RW * FZ_TX;   // 128 pixels per tile = N of the MMA
constexpr int FZ_NPOS = 16;                                    // low-res positions per strip = K of the MMA
constexpr int FZ_NSTG = 16, FZ_INFL = 4;                       // strip ring slots / strips in flight per loader thread
constexpr int FZ_EPI_WARPS = 8, FZ_LD_WARPS = 2, FZ_BB_WARPS = 8;   // 19 warps: the same 640-thread register allocation as 17
constexpr int FZ_THREADS = 32 * (FZ_EPI_WARPS + 1 + FZ_LD_WARPS + FZ_BB_WARPS);
constexpr int FZ_LD_T0 = 32 * (FZ_EPI_WARPS + 1), FZ_BB_T0 = FZ_LD_T0 + 32 * FZ_LD_WARPS;

__device__ __forceinline__ int fz_reflect(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}
__device__ __forceinline__ void fz_cubic(float t, float (&c)[4]) {   // torch upsample_bicubic2d, A = -0.75
  const float A = -0.75f;
  const float x0 = t + 1.f, x1 = t, x2 = 1.f - t, x3 = 2.f - t;
  c[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  c[1] = ((A + 2.f) * x1 - (A + 3.f)) * x1 * x1 + 1.f;
  c[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  c[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}
__device__ __forceinline__ void fz_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fz_mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t fz_pack(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Tables (fp16): tab(c, d, t) = sum over the bicubic taps b of high-res coordinate q = reflect(c + t - R) that land
// on low-res coordinate (c >> 1) + d - DO of coeff_b(q); reflect padding and border clamps are resolved here.
// Stored in mma.sync fragment order, one uint4 per (coordinate, lane):
//   columns (n = W2): A operand of stage A, TX^T[dx][j]:  {(g, 2tig), (g+8, 2tig), (g, 2tig+8), (g+8, 2tig+8)}
//   rows    (n = H2): B operand of stage B, TY^T[dy][i]:  {(g, 2tig), (g, 2tig+8), (g+8, 2tig), (g+8, 2tig+8)}
// where (row, col) = (d, t) and each register packs columns col, col+1.
__device__ __forceinline__ float fz_tab_entry(int c, int d, int t, int n, int R) {
  const int D = 2 * R + 1, DO = (R + 3) / 2, nl = n >> 1;
  if (t >= D) return 0.f;
  const int q = fz_reflect(c + t - R, n), u = q >> 1, odd = q & 1;
  float cf[4];
  fz_cubic(odd ? 0.25f : 0.75f, cf);                   // even q = 2u: taps u-2..u+1; odd: u-1..u+2
  float acc = 0.f;
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const int l = min(max(u - 2 + odd + b, 0), nl - 1);
    if (l - (c >> 1) + DO == d) acc += cf[b];
  }
  return acc;
}
__global__ void fz_tables_kernel(int W2, int H2, int R, uint4* __restrict__ tabx, uint4* __restrict__ taby) {
  pdl_grid_sync();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (W2 + H2) * 32) return;
  const int lane = idx & 31, cc = idx >> 5, g = lane >> 2, tig = lane & 3;
  const bool rows = cc >= W2;
  const int c = rows ? cc - W2 : cc, n = rows ? H2 : W2;
  uint32_t r[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int hi_row = rows ? (k >> 1) : (k & 1), hi_col = rows ? (k & 1) : (k >> 1);
    const int d = g + hi_row * 8, t = 2 * tig + hi_col * 8;
    r[k] = fz_pack(fz_tab_entry(c, d, t, n, R), fz_tab_entry(c, d, t + 1, n, R));
  }
  (rows ? taby : tabx)[(size_t)c * 32 + lane] = make_uint4(r[0], r[1], r[2], r[3]);
}

// Composite kernels K'^T[dx][dy] of every output pixel, one warp per pixel (grid-stride):
//   stage A  T^T[dx][i]  = sum_j TX^T[x][dx][j] * k[i][j]        (mma.sync fp16, A = table fragment, B = the weight row)
//   stage B  K'^T[dx][dy] = sum_i T^T[dx][i] * TY^T[y][dy][i]     (the accumulators of stage A are the A fragment)
// kc[p][dy * (2 DO + 1) + dx] (bf16, row stride 128) is what the banded GEMM scatters into its B tiles.
// MODE 0: every pixel of n crops (dense, kern / kc indexed [crop][y][x]).
// MODE 1: every pixel of the image-level stage canvas (ih x iw): kern_img -> kc_img with the bicubic tables of a
//         crop-INTERIOR pixel of the same parity (the tables only depend on the parity there: the composite kernel of an
//         interior pixel is the same for every crop that contains it, jbu_share.cuh).
// MODE 2: the border frames (fb = CSEG_JBU_FB_COMP) of n crops, compact kc; the fixed-up kernel of a frame pixel comes
//         from the per-crop border tensor (frame CSEG_JBU_FB_RANGE) or, further inside, from the image-level tensor.
template <int R, int MODE>
__global__ void __launch_bounds__(256, MODE == 2 ? 6 : 1) fz_composite_kernel(const bf16* __restrict__ kern, int ldk, long long n_px, int H2,
                                                           int W2, const uint4* __restrict__ tabx,
                                                           const uint4* __restrict__ taby, bf16* __restrict__ kc,
                                                           const bf16* __restrict__ kern_img, const ShareGeom sg, int iw) {
  pdl_grid_sync();
  constexpr int D = 2 * R + 1, DO = (R + 3) / 2, DC = 2 * DO + 1;
  const int lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
  int boff[2][2][2];                                           // [n-block of i][k half of j][element]; -1 = padding
#pragma unroll
  for (int nb = 0; nb < 2; ++nb)
#pragma unroll
    for (int kh = 0; kh < 2; ++kh)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = nb * 8 + g, j = kh * 8 + 2 * tig + e;
        boff[nb][kh][e] = (i < D && j < D) ? i * D + j : -1;
      }
  // every warp takes a contiguous run of pixels, so (x, y) advance incrementally (no per-pixel divisions)
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long per = (n_px + nwarps - 1) / nwarps, p_begin = warp0 * per, p_end = min(n_px, p_begin + per);
  const int roww = MODE == 1 ? iw : W2;
  const int rb_c = border_rows(H2, W2, CSEG_JBU_FB_COMP), rb_r = border_rows(H2, W2, CSEG_JBU_FB_RANGE);
  int x = (int)(p_begin % roww), y = MODE == 1 ? (int)(p_begin / roww) : (int)((p_begin / roww) % H2);
  // MODE 2: (crop, compact frame row r) -> (by, bx) also advance incrementally: the divisions of border_coords and the
  // window look-up happen once per run / per crop, not once per pixel
  constexpr int FBC = CSEG_JBU_FB_COMP;
  const int s1 = FBC * W2, s2 = 2 * FBC * W2;
  int crop = 0, r = 0, bx = 0, by = 0;
  size_t org0 = 0;
  if (MODE == 2 && p_begin < p_end) {
    crop = (int)(p_begin / rb_c);
    r = (int)(p_begin - (long long)crop * rb_c);
    border_coords(r, H2, W2, FBC, by, bx);
    org0 = (size_t)(sg.wins[crop * 4] >> sg.shift) * sg.pitch + (sg.wins[crop * 4 + 1] >> sg.shift);
  }
  for (long long p = p_begin; p < p_end; ++p, ++x) {
    if (MODE != 2 && x == roww) {
      x = 0;
      if (++y == H2 && MODE == 0) y = 0;
    }
    int tx = x, ty = y;
    const bf16* krow;
    if (MODE == 0) {
      krow = kern + p * ldk;
    } else if (MODE == 1) {
      tx = 16 + (x & 1);
      ty = 16 + (y & 1);
      krow = kern + p * ldk;
    } else {
      tx = bx;
      ty = by;
      if (border_interior(ty, tx, H2, W2, CSEG_JBU_FB_RANGE)) {
        krow = kern_img + (org0 + (size_t)ty * sg.pitch + tx) * ldk;
      } else {
        krow = kern + ((size_t)crop * rb_r + border_index(ty, tx, H2, W2, CSEG_JBU_FB_RANGE)) * ldk;
      }
      // next frame pixel (layout of jbu_share.cuh: top strip, bottom strip, then 32-pixel groups of the rows in between)
      if (++r == rb_c) {
        r = 0;
        ++crop;
        bx = by = 0;
        if (p + 1 < p_end) org0 = (size_t)(sg.wins[crop * 4] >> sg.shift) * sg.pitch + (sg.wins[crop * 4 + 1] >> sg.shift);
      } else if (r < s2) {
        if (++bx == W2) {
          bx = 0;
          ++by;
          if (r == s1) by = H2 - FBC;
        }
      } else {
        const int q = r - s2, c = q & 31;
        by = FBC + (q >> 5);
        bx = c < 16 ? c : W2 - 32 + c;
      }
    }
    const uint4 txv = __ldg(tabx + (size_t)tx * 32 + lane), tyv = __ldg(taby + (size_t)ty * 32 + lane);
    const uint32_t ta[4] = {txv.x, txv.y, txv.z, txv.w};
    const uint32_t tyb[2][2] = {{tyv.x, tyv.y}, {tyv.z, tyv.w}};
    const unsigned short* wrow = reinterpret_cast<const unsigned short*>(krow);
    float tacc[2][4];
#pragma unroll
    for (int nb = 0; nb < 2; ++nb) {
      uint32_t kb[2];
#pragma unroll
      for (int kh = 0; kh < 2; ++kh) {
        const int o0 = boff[nb][kh][0], o1 = boff[nb][kh][1];
        const uint32_t lo = o0 >= 0 ? (uint32_t)__ldg(wrow + o0) : 0u, hi = o1 >= 0 ? (uint32_t)__ldg(wrow + o1) : 0u;
        kb[kh] = fz_pack(__uint_as_float(lo << 16), __uint_as_float(hi << 16));
      }
      tacc[nb][0] = tacc[nb][1] = tacc[nb][2] = tacc[nb][3] = 0.f;
      fz_mma_f16(tacc[nb], ta, kb[0], kb[1]);
    }
    const uint32_t a2[4] = {fz_pack(tacc[0][0], tacc[0][1]), fz_pack(tacc[0][2], tacc[0][3]),
                            fz_pack(tacc[1][0], tacc[1][1]), fz_pack(tacc[1][2], tacc[1][3])};
    bf16* orow = kc + p * 128;
#pragma unroll
    for (int nb = 0; nb < 2; ++nb) {
      float kacc[4] = {0.f, 0.f, 0.f, 0.f};
      fz_mma_f16(kacc, a2, tyb[nb][0], tyb[nb][1]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int dx = g + (e >> 1) * 8, dy = nb * 8 + 2 * tig + (e & 1);
        if (dx < DC && dy < DC) orow[dy * DC + dx] = __float2bfloat16_rn(kacc[e]);
      }
    }
  }
}

__device__ __forceinline__ uint64_t fz_adesc(uint32_t saddr, uint32_t lbo_bytes) {   // MN-major SWIZZLE_128B (jbu_apply_tc.cu)
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t fz_idesc() {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((uint32_t)(FZ_NPX >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

template <int R, int MH>
struct FzCfg {
  static constexpr int DO = (R + 3) / 2;                         // composite kernel spans 2 DO + 1 low-res taps
  static constexpr int NLR = (FZ_RW + 2 * R) / 2 + 3;             // low-res rows per tile (10 / 8)
  static constexpr int NQUAD = (NLR + 3) / 4;                     // four strip rows share one 128 B band-tile row
  static constexpr int CH = 128 * MH;
  static constexpr int CHUNK_BYTES = FZ_NPOS * 128;               // one 64-channel chunk of a strip (16 rows x 128 B)
  static constexpr int A_STAGE = (CH / 64) * CHUNK_BYTES;         // 4 / 8 KB
  static constexpr int B_QUAD = FZ_NPX * 128, B_BUF = NQUAD * B_QUAD;
  static constexpr int A_OFF = 0, B_OFF = FZ_NSTG * A_STAGE, BAR_OFF = B_OFF + 2 * B_BUF;
  static constexpr int NBARS = 2 * FZ_NSTG + 8;
  static constexpr int SMEM_BYTES = BAR_OFF + NBARS * 8 + 16 + 1024;
  static constexpr int TMEM_COLS = (2 * MH * FZ_NPX <= 128) ? 128 : ((2 * MH * FZ_NPX <= 256) ? 256 : 512);
};

// CST: compile-time channel count of the output rows (0 = runtime C): the 16 two-byte stores of an output row then take
// their offsets m * C * 2 as immediates of ONE base address instead of a 64-bit add each.
template <int R, int MH, int CST>
__global__ void __launch_bounds__(FZ_THREADS, 1)
jbu_apply_fused_kernel(const bf16* __restrict__ src, int h, int w, int C, const bf16* __restrict__ kc,
                       bf16* __restrict__ dst, int nx, int ny, int nslab, int seg, int nseg, int total_units,
                       const bf16* __restrict__ kc_img, const ShareGeom sg, int diag) {
  using Cf = FzCfg<R, MH>;
  constexpr int DO = Cf::DO, NLR = Cf::NLR;
  const int H2 = 2 * h, W2 = 2 * w;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic on the __shared__ array: accesses compile to LDS / STS, not generic LD / ST
  uint64_t* bars = (uint64_t*)(smem + Cf::BAR_OFF);
  uint32_t* tmem_slot = (uint32_t*)(bars + Cf::NBARS);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t a_full0 = smem_u32(bars), a_empty0 = a_full0 + FZ_NSTG * 8;
  const uint32_t b_full0 = a_empty0 + FZ_NSTG * 8, b_empty0 = b_full0 + 16;
  const uint32_t t_full0 = b_empty0 + 16, t_empty0 = t_full0 + 16;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < FZ_NSTG; ++i) {
      mbar_init(a_full0 + i * 8, 32 * FZ_LD_WARPS);
      mbar_init(a_empty0 + i * 8, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(b_full0 + i * 8, 32 * FZ_BB_WARPS);
      mbar_init(b_empty0 + i * 8, 1);
      mbar_init(t_full0 + i * 8, 1);
      mbar_init(t_empty0 + i * 8, FZ_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == FZ_EPI_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)Cf::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_grid_sync();   // everything above is input-independent and overlaps the tail of the previous kernel
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // unit -> column (x tile, crop, channel slab) and the run of y tiles [yt0, yt1)
  auto unit_coords = [&](int unit, int& x0, int& crop, int& c0, int& yt0, int& yt1) {
    const int col = unit / nseg, sgi = unit - col * nseg;
    const int xt = col % nx, z = col / nx;
    x0 = xt * FZ_TX;
    crop = z / nslab;
    c0 = (z % nslab) * Cf::CH;
    yt0 = sgi * seg;
    yt1 = min(ny, yt0 + seg);
  };

  if (warp < FZ_EPI_WARPS) {
    // ---------------- epilogue: lane = channel, 64 pixel columns per channel slab ----------------
    uint32_t tl = 0;
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x)
    {
     int x0, crop, c0, yt0, yt1;
     unit_coords(unit, x0, crop, c0, yt0, yt1);
     for (int yt = yt0; yt < yt1; ++yt, ++tl) {
      const int y0 = yt * FZ_RW;
      const uint32_t as = tl & 1, aph = (tl >> 1) & 1;
      mbar_wait(t_full0 + as * 8, aph);
      tc_fence_after();
      // 8 warps: TMEM lane quadrant = warp % 4 (32 channels), warp / 4 selects one half of the MH * 4 column steps (with two
      // channel slabs: one slab each) -- the epilogue was the busiest role of the kernel (94 % of its warps' samples)
      constexpr int HSTEPS = MH * (FZ_NPX / 32) / (FZ_EPI_WARPS / 4);
      const int wq = warp & 3, wh = warp >> 2;
#pragma unroll 1
      for (int hc = wh * HSTEPS; hc < (wh + 1) * HSTEPS; ++hc) {  // 32 pixel columns (two output rows) per step
        const int half = hc / (FZ_NPX / 32), cb = hc % (FZ_NPX / 32);
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)((as * MH + half) * FZ_NPX + cb * 32), r);
        const int Cs = CST ? CST : C;
        bf16* obase = dst + (size_t)crop * H2 * W2 * Cs + c0 + half * 128 + wq * 32 + lane;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int y = y0 + cb * 2 + rr;
          if (y >= H2) continue;
          bf16* o = obase + ((size_t)y * W2 + x0) * Cs;
          if (diag & 4) continue;                                  // diagnostic: no output stores
          const int mmax = min(16, W2 - x0);
          if (mmax == 16) {
#pragma unroll
            for (int m = 0; m < 16; ++m) o[(size_t)m * Cs] = __float2bfloat16_rn(__uint_as_float(r[rr * 16 + m]));
          } else {
#pragma unroll
            for (int m = 0; m < 16; ++m)
              if (m < mmax) o[(size_t)m * Cs] = __float2bfloat16_rn(__uint_as_float(r[rr * 16 + m]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_empty0 + as * 8);
     }
    }
  } else if (warp == FZ_EPI_WARPS) {
    // ---------------- MMA issuer: one MMA per low-res row and channel half ----------------
    // the whole warp runs the loop (warp-uniform control flow); one elected lane issues the MMAs and commits
    {
      constexpr uint32_t idesc = fz_idesc();
      uint32_t tl = 0, base_it = 0, ready = 0;                  // base_it: ring fill counter of the unit's first row
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        int x0, crop, c0, yt0, yt1;
        unit_coords(unit, x0, crop, c0, yt0, yt1);
        const int ntile = yt1 - yt0;
        for (int i = 0; i < ntile; ++i, ++tl) {
          const uint32_t as = tl & 1, aph = (tl >> 1) & 1;
          mbar_wait(t_empty0 + as * 8, aph ^ 1);
          mbar_wait(b_full0 + as * 8, aph);
          tc_fence_after();
          const uint32_t bbuf = smem_base + Cf::B_OFF + as * Cf::B_BUF;
          const bool last = (i == ntile - 1);
          for (int s = 0; s < NLR; ++s) {
            const uint32_t g = base_it + (uint32_t)((FZ_RW / 2) * i + s);   // row s of this tile = row (RW/2) i + s of the run
            const uint32_t st = g % FZ_NSTG, ph = (g / FZ_NSTG) & 1;
            if (g >= ready) {                                      // strips land in order: rows below `ready` are known to
              mbar_wait(a_full0 + st * 8, ph);                     // be resident (NLR - RW/2 rows of every tile after the
              tc_fence_after();                                    // first one of a run were waited for by the tile above)
              ready = g + 1;
            }
            const uint32_t a_src = smem_base + Cf::A_OFF + st * Cf::A_STAGE;
            const uint64_t bdesc = make_sdesc(bbuf + (s >> 2) * Cf::B_QUAD + (s & 3) * 32);
            if (elect_one()) {
#pragma unroll
              for (int half = 0; half < MH; ++half) {
                const uint64_t adesc = fz_adesc(a_src + half * 2 * Cf::CHUNK_BYTES, Cf::CHUNK_BYTES);
                if (!(diag & 16))                                  // diagnostic bit 16: no MMAs (commits still fire)
                  umma_f16(tmem_base + (uint32_t)((as * MH + half) * FZ_NPX), adesc, bdesc, idesc, s > 0 ? 1u : 0u);
              }
              // the two top rows leave the window of the next tile (all rows after the last tile of the run)
              if (s < FZ_RW / 2 || last) umma_commit(a_empty0 + st * 8);
              if (s == NLR - 1) {
                umma_commit(t_full0 + as * 8);
                umma_commit(b_empty0 + as * 8);
              }
            }
            __syncwarp();
          }
        }
        base_it += (uint32_t)(NLR + (FZ_RW / 2) * (ntile - 1));
      }
    }
  } else if (warp < FZ_EPI_WARPS + 1 + FZ_LD_WARPS) {
    // ---------------- strip loaders: NLR low-res rows x 16 positions x CH channels per tile ----------------
    const int lt = tid - FZ_LD_T0;
    constexpr int CPR = Cf::CH / 8;                              // 16-byte chunks per position
    constexpr int LPT = FZ_NPOS * CPR / (32 * FZ_LD_WARPS);      // chunks per thread per strip (2 / 4)
    uint32_t it = 0;
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
      int x0, crop, c0, yt0, yt1;
      unit_coords(unit, x0, crop, c0, yt0, yt1);
      const bf16* sc = src + (size_t)crop * h * w * C + c0;
      const int lx0 = (x0 >> 1) - DO, ly0 = ((yt0 * FZ_RW) >> 1) - DO;
      const int nrows = NLR + (FZ_RW / 2) * (yt1 - yt0 - 1);     // low-res rows of the whole run
      int src_off[LPT];
      uint32_t dst_off[LPT];
#pragma unroll
      for (int k = 0; k < LPT; ++k) {
        const int e = lt + k * 32 * FZ_LD_WARPS;
        const int p = e / CPR, c8 = e % CPR;
        const int xx = min(max(lx0 + p, 0), w - 1);              // outside the image the band weights are zero
        src_off[k] = xx * C + c8 * 8;
        dst_off[k] = (uint32_t)((c8 >> 3) * Cf::CHUNK_BYTES + p * 128 + (((c8 & 7) ^ (p & 7)) << 4));
      }
      for (int s = 0; s < nrows; ++s, ++it) {
        const uint32_t st = it % FZ_NSTG, ph = (it / FZ_NSTG) & 1;
        mbar_wait(a_empty0 + st * 8, ph ^ 1);
        const int yy = min(max(ly0 + s, 0), h - 1);
        const bf16* rowp = sc + (size_t)yy * w * C;
        const uint32_t base = smem_base + Cf::A_OFF + st * Cf::A_STAGE;
        if (!(diag & 2)) {                                       // diagnostic bit 2: no strip fetches
#pragma unroll
          for (int k = 0; k < LPT; ++k)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(base + dst_off[k]), "l"(rowp + src_off[k]) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (it >= FZ_INFL - 1) {
          asm volatile("cp.async.wait_group %0;" ::"n"(FZ_INFL - 1) : "memory");
          fz_fence_proxy_async();
          mbar_arrive(a_full0 + ((it - (FZ_INFL - 1)) % FZ_NSTG) * 8);
        }
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    fz_fence_proxy_async();
    for (uint32_t j = (it >= FZ_INFL - 1 ? it - (FZ_INFL - 1) : 0); j < it; ++j) mbar_arrive(a_full0 + (j % FZ_NSTG) * 8);
  } else {
    // ---------------- band builders: scatter the composite kernels of the tile's 128 pixels ----------------
    // A thread owns ONE pixel of the tile (n = bt % 128) and one half of its composite kernel (the first NC0 or the last
    // NC1 16-byte chunks; the half is warp-uniform).  Where a weight lands in the band tiles depends on (n, tap) only, not
    // on the tile: the swizzled offsets are computed once per kernel and kept as 16-bit pairs in registers, so a tile costs
    // one pointer computation, NC0 16-byte loads and one 2-byte store (+ ~3 integer instructions) per weight.
    const int bt = tid - FZ_BB_T0;                               // 0..255
    constexpr int DC = 2 * DO + 1, NCH = (DC * DC + 7) / 8;      // 16-byte chunks per composite kernel (11 / 7)
    constexpr int NC0 = (NCH + 1) / 2, NC1 = NCH - NC0;
    static_assert(32 * FZ_BB_WARPS == 2 * FZ_NPX, "two builder threads per pixel");
    static_assert(Cf::B_BUF <= 65536, "band-tile offsets are kept in 16 bits");
    const int n = bt & (FZ_NPX - 1), half = bt >> 7;
    const int nck = half ? NC1 : NC0, v0 = half ? NC0 : 0;
    uint32_t offp[NC0 * 4];
    {
      const int r = n >> 4, m = n & 15;
      const int sr0 = r >> 1, kk0 = m >> 1;                      // (y >> 1) - (y0 >> 1), (x >> 1) - (x0 >> 1)
      const uint32_t rowb = (uint32_t)((n >> 3) * 1024 + (n & 7) * 128);
#pragma unroll
      for (int k = 0; k < NC0; ++k)
#pragma unroll
        for (int q2 = 0; q2 < 4; ++q2) {
          uint32_t o[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int t = (v0 + k) * 8 + 2 * q2 + e;
            const int dy = t / DC, dx = t - dy * DC;
            const int sr = sr0 + dy, col = (sr & 3) * 16 + kk0 + dx;
            o[e] = (t < DC * DC) ? rowb + (uint32_t)((sr >> 2) * Cf::B_QUAD) + (uint32_t)((((col >> 3) ^ (n & 7)) << 4) | ((col & 7) << 1)) : 0u;
          }
          offp[k * 4 + q2] = o[0] | (o[1] << 16);
        }
    }
    // The weights of tile i + 1 are requested before tile i is scattered (registers, one tile ahead): the composite kernels
    // stream from HBM (616 MB per 48 crops at the last stage) and a tile's loads would otherwise expose a full DRAM round
    // trip between two scatters.
    int unit = blockIdx.x, yt = 0, yt1 = 0, x0 = 0, crop = 0, c0 = 0;
    size_t org = 0, cbase = 0;
    bool have = unit < total_units;
    auto enter_unit = [&]() {
      int yt0;
      unit_coords(unit, x0, crop, c0, yt0, yt1);
      yt = yt0;
      // shared kernels (jbu_share.cuh): interior pixels read the image-level composite kernels at the crop's origin,
      // border-frame pixels the crop's compact frame tensor
      org = 0;
      cbase = (size_t)crop * H2 * W2;
      if (kc_img != nullptr) {
        org = (size_t)(sg.wins[crop * 4] >> sg.shift) * sg.pitch + (sg.wins[crop * 4 + 1] >> sg.shift);
        cbase = (size_t)crop * border_rows(H2, W2, CSEG_JBU_FB_COMP);
      }
    };
    auto locate = [&]() -> const uint4* {                        // composite weights of this thread's pixel in tile (unit, yt)
      const int y = yt * FZ_RW + (n >> 4), x = x0 + (n & 15);
      if (y >= H2 || x >= W2 || (diag & 8)) return nullptr;      // outside the image: zero weights (diagnostic bit 8: no loads)
      const bf16* kp;
      if (kc_img == nullptr) kp = kc + (cbase + (size_t)y * W2 + x) * 128;
      else if (border_interior(y, x, H2, W2, CSEG_JBU_FB_COMP)) kp = kc_img + (org + (size_t)y * sg.pitch + x) * 128;
      else kp = kc + (cbase + border_index(y, x, H2, W2, CSEG_JBU_FB_COMP)) * 128;
      return reinterpret_cast<const uint4*>(kp) + v0;
    };
    uint4 wn[NC0];
#pragma unroll
    for (int k = 0; k < NC0; ++k) wn[k] = make_uint4(0, 0, 0, 0);
    if (have) {
      enter_unit();
      const uint4* kv = locate();
      if (kv != nullptr) {
#pragma unroll
        for (int k = 0; k < NC0; ++k)
          if (k < nck) wn[k] = __ldg(kv + k);
      }
    }
    for (uint32_t tl = 0; have; ++tl) {
      const uint32_t as = tl & 1, aph = (tl >> 1) & 1;
      if (++yt >= yt1) {                                         // next tile of the run, or the first tile of the next unit
        unit += gridDim.x;
        have = unit < total_units;
        if (have) enter_unit();
      }
      const uint4* kvn = have ? locate() : nullptr;              // weights of the NEXT tile: requested chunk by chunk as the
      mbar_wait(b_empty0 + as * 8, aph ^ 1);                     // registers of the current tile are scattered (one tile ahead)
      const uint32_t bb = smem_base + Cf::B_OFF + as * Cf::B_BUF;
      if (tl < 2) {                                              // zero background, once per buffer
        uint8_t* bbuf = smem + Cf::B_OFF + as * Cf::B_BUF;
        for (int e = bt; e < Cf::B_BUF / 16; e += 32 * FZ_BB_WARPS) reinterpret_cast<uint4*>(bbuf)[e] = make_uint4(0, 0, 0, 0);
        asm volatile("bar.sync 2, %0;" ::"n"(32 * FZ_BB_WARPS) : "memory");
      }
#pragma unroll
      for (int k = 0; k < NC0; ++k) {
        if (k >= nck) break;                                     // warp-uniform
        if (!(diag & 1)) {                                       // diagnostic bit 1: no band scatter
          const uint32_t w4[4] = {wn[k].x, wn[k].y, wn[k].z, wn[k].w};
#pragma unroll
          for (int q2 = 0; q2 < 4; ++q2) {
            const uint32_t op = offp[k * 4 + q2];
            // taps beyond DC * DC only occur in the last chunk of the second half
            if ((NC0 + k) * 8 + 2 * q2 < DC * DC || !half)
              asm volatile("st.shared.u16 [%0], %1;" ::"r"(bb + (op & 0xFFFFu)), "h"((unsigned short)(w4[q2] & 0xFFFFu)) : "memory");
            if ((NC0 + k) * 8 + 2 * q2 + 1 < DC * DC || !half)
              asm volatile("st.shared.u16 [%0], %1;" ::"r"(bb + (op >> 16)), "h"((unsigned short)(w4[q2] >> 16)) : "memory");
          }
        }
        wn[k] = (kvn != nullptr) ? __ldg(kvn + k) : make_uint4(0, 0, 0, 0);
      }
      fz_fence_proxy_async();
      mbar_arrive(b_full0 + as * 8);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == FZ_EPI_WARPS) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cf::TMEM_COLS) : "memory");
  }
}

template <int R, int MH, int CST>
int launch_apply_kernel_c(const bf16* src, int n_crops, int h, int w, int C, const bf16* kc, bf16* dst, const bf16* kc_img,
                          const ShareGeom& sg, cudaStream_t st) {
  using Cf = FzCfg<R, MH>;
  const int H2 = 2 * h, W2 = 2 * w;
  CSEG_SET_SMEM((jbu_apply_fused_kernel<R, MH, CST>), Cf::SMEM_BYTES);
  const int nx = cdiv(W2, FZ_TX), ny = cdiv(H2, FZ_RW), nslab = C / Cf::CH;
  const long long ncols = (long long)nx * n_crops * nslab;
  CSEG_REQUIRE(ncols * ny < (1ll << 31), "jbu_apply(bf16): too many tiles");
  // tiles per run: longer runs re-use more source rows (a run of s tiles fetches NLR + 2 (s - 1) strips instead of
  // NLR s) but leave fewer, coarser units to balance over the SMs.  Cost of a unit in strip-sized L2 transfers: its
  // strips + 6 per tile (output + composite kernels); pick the run length with the cheapest slowest SM.
  int seg = 1;
  {
    const int sms = sm_count();
    double best = 1e30;
    for (int s = 1; s <= std::min(ny, 16); ++s) {
      const long long units = ncols * cdiv(ny, s);
      const double cost = (double)cdiv(units, sms) * (Cf::NLR + (FZ_RW / 2) * (s - 1) + 12.0 * s);
      if (cost < best - 1e-9) { best = cost; seg = s; }
    }
    if (const char* e = getenv("CSEG_APPLY_SEG")) seg = std::max(1, std::min(atoi(e), ny));   // A/B measurements
  }
  const int nseg = cdiv(ny, seg);
  const long long total = ncols * nseg;
  const int grid = (int)std::min<long long>(total, sm_count());
  static int diag = -1;                            // CSEG_APPLY_DIAG: role knock-out bits for timing experiments (results invalid)
  if (diag < 0) {
    const char* e = getenv("CSEG_APPLY_DIAG");
    diag = e ? atoi(e) : 0;
  }
  cseg_launch(jbu_apply_fused_kernel<R, MH, CST>, dim3(grid), dim3(FZ_THREADS), Cf::SMEM_BYTES, st, src, h, w, C, kc,
              dst, nx, ny, nslab, seg, nseg, (int)total, kc_img, sg, diag);
  CSEG_LAUNCH_CHECK("jbu_apply_fused");
  return 0;
}
template <int R, int MH>
int launch_apply_kernel(const bf16* src, int n_crops, int h, int w, int C, const bf16* kc, bf16* dst, const bf16* kc_img,
                        const ShareGeom& sg, cudaStream_t st) {
  if (C == 256) return launch_apply_kernel_c<R, MH, 256>(src, n_crops, h, w, C, kc, dst, kc_img, sg, st);   // basis form (ViT-B/L)
  if (C == 512) return launch_apply_kernel_c<R, MH, 512>(src, n_crops, h, w, C, kc, dst, kc_img, sg, st);   // literal form
  return launch_apply_kernel_c<R, MH, 0>(src, n_crops, h, w, C, kc, dst, kc_img, sg, st);
}

template <int R, int MH>
int launch_fused(const bf16* src, int n_crops, int h, int w, int C, const bf16* kern, int ldk, bf16* dst, uint8_t* scratch,
                 cudaStream_t st) {
  const int H2 = 2 * h, W2 = 2 * w;
  const long long n_px = (long long)n_crops * H2 * W2;
  bf16* kc = reinterpret_cast<bf16*>(scratch);                       // composite kernels [n_px, 128]
  uint4* tabx = reinterpret_cast<uint4*>(scratch + n_px * 256);      // then the bicubic tables
  uint4* taby = tabx + (size_t)W2 * 32;
  cseg_launch(fz_tables_kernel, dim3(cdiv((W2 + H2) * 32, 256)), dim3(256), 0, st, W2, H2, R, tabx, taby);
  CSEG_LAUNCH_CHECK("jbu_apply_tables");
  const int cblocks = (int)std::min<long long>(cdiv(n_px, 8), (long long)sm_count() * 8);
  const ShareGeom none = {nullptr, 0, 0};
  cseg_launch(fz_composite_kernel<R, 0>, dim3(cblocks), dim3(256), 0, st, kern, ldk, n_px, H2, W2, (const uint4*)tabx,
              (const uint4*)taby, kc, (const bf16*)nullptr, none, 0);
  CSEG_LAUNCH_CHECK("jbu_apply_composite");
  return launch_apply_kernel<R, MH>(src, n_crops, h, w, C, kc, dst, nullptr, none, st);
}

// image-level composite kernels (MODE 1); tabs: (gh + gw) * 512 bytes
template <int R>
int launch_composite_image(const bf16* kern_img, int ldk, int ih, int iw, int gh, int gw, bf16* kc_img, uint8_t* tabs,
                           cudaStream_t st) {
  uint4* tabx = reinterpret_cast<uint4*>(tabs);
  uint4* taby = tabx + (size_t)gw * 32;
  cseg_launch(fz_tables_kernel, dim3(cdiv((gw + gh) * 32, 256)), dim3(256), 0, st, gw, gh, R, tabx, taby);
  CSEG_LAUNCH_CHECK("jbu_apply_tables");
  const long long n_px = (long long)ih * iw;
  const int cblocks = (int)std::min<long long>(cdiv(n_px, 8), (long long)sm_count() * 8);
  const ShareGeom none = {nullptr, 0, 0};
  cseg_launch(fz_composite_kernel<R, 1>, dim3(cblocks), dim3(256), 0, st, kern_img, ldk, n_px, gh, gw, (const uint4*)tabx,
              (const uint4*)taby, kc_img, (const bf16*)nullptr, none, iw);
  CSEG_LAUNCH_CHECK("jbu_composite_image");
  return 0;
}

// shared form of launch_fused: border-frame composites (MODE 2) into `scratch`, then the banded GEMM with indirection
template <int R, int MH>
int launch_fused_shared(const bf16* src, int n_crops, int h, int w, int C, const bf16* kern_border, const bf16* kern_img,
                        const bf16* kc_img, const ShareGeom& sg, int ldk, bf16* dst, uint8_t* scratch, cudaStream_t st) {
  const int H2 = 2 * h, W2 = 2 * w;
  const long long n_px = (long long)n_crops * border_rows(H2, W2, CSEG_JBU_FB_COMP);
  bf16* kc = reinterpret_cast<bf16*>(scratch);                       // compact frame composites [n_px, 128]
  uint4* tabx = reinterpret_cast<uint4*>(scratch + n_px * 256);
  uint4* taby = tabx + (size_t)W2 * 32;
  cseg_launch(fz_tables_kernel, dim3(cdiv((W2 + H2) * 32, 256)), dim3(256), 0, st, W2, H2, R, tabx, taby);
  CSEG_LAUNCH_CHECK("jbu_apply_tables");
  const int cblocks = (int)std::min<long long>(cdiv(n_px, 8), (long long)sm_count() * 8);
  cseg_launch(fz_composite_kernel<R, 2>, dim3(cblocks), dim3(256), 0, st, kern_border, ldk, n_px, H2, W2, (const uint4*)tabx,
              (const uint4*)taby, kc, kern_img, sg, 0);
  CSEG_LAUNCH_CHECK("jbu_apply_composite_border");
  return launch_apply_kernel<R, MH>(src, n_crops, h, w, C, kc, dst, kc_img, sg, st);
}

}  // namespace

// returns 1 when the shape is not covered (caller runs bicubic2x + the stand-alone adaptive conv instead).
// scratch: the caller's hr_scratch (n * 2h * 2w * C elements >= what is used here): composite kernels
// [n * 2h * 2w, 128] bf16, then (2h + 2w) * 512 bytes of bicubic tables.
int cseg_jbu_apply_fused(const bf16* src, int n_crops, int h, int w, int C, const bf16* kern, int ldk, int radius,
                         bf16* dst, void* scratch, cudaStream_t st) {
  if (C % 128 != 0 || ldk % 8 != 0 || (radius != 5 && radius != 3)) return 1;
  if (h < radius + 2 || w < 12 || scratch == nullptr) return 1;          // single reflection; strips of 16 positions
  if ((long long)C * 2 < 256 + 16) return 1;                             // scratch holds 256 B per pixel + the tables
  if (((uintptr_t)src & 15) != 0 || ((uintptr_t)kern & 15) != 0 || ((uintptr_t)scratch & 15) != 0) return 1;
  uint8_t* tabs = (uint8_t*)scratch;
  if (C % 256 == 0) {
    if (radius == 5) return launch_fused<5, 2>(src, n_crops, h, w, C, kern, ldk, dst, tabs, st);
    return launch_fused<3, 2>(src, n_crops, h, w, C, kern, ldk, dst, tabs, st);
  }
  if (radius == 5) return launch_fused<5, 1>(src, n_crops, h, w, C, kern, ldk, dst, tabs, st);
  return launch_fused<3, 1>(src, n_crops, h, w, C, kern, ldk, dst, tabs, st);
}

// image-level composite kernels; returns 1 when the shape is not covered
int cseg_jbu_composite_image_tc(const bf16* kern_img, int ldk, int ih, int iw, int gh, int gw, int radius, bf16* kc_img,
                                void* tabs, cudaStream_t st) {
  if (ldk % 8 != 0 || (radius != 5 && radius != 3)) return 1;
  if (((uintptr_t)kern_img & 15) != 0 || ((uintptr_t)kc_img & 15) != 0 || ((uintptr_t)tabs & 15) != 0) return 1;
  if (radius == 5) return launch_composite_image<5>(kern_img, ldk, ih, iw, gh, gw, kc_img, (uint8_t*)tabs, st);
  return launch_composite_image<3>(kern_img, ldk, ih, iw, gh, gw, kc_img, (uint8_t*)tabs, st);
}

// cseg_jbu_apply with kernels shared across crops (include/clipseg.h); returns 1 when the shape is not covered
int cseg_jbu_apply_shared_tc(const bf16* src, int n_crops, int h, int w, int C, const bf16* kern_border, const bf16* kern_img,
                             const bf16* kc_img, const ShareGeom& sg, int ldk, int radius, bf16* dst, void* scratch,
                             cudaStream_t st) {
  if (C % 128 != 0 || ldk % 8 != 0 || (radius != 5 && radius != 3)) return 1;
  if (2 * h < 2 * CSEG_JBU_FB_COMP + 2 || 2 * w < 34) return 1;
  if (((uintptr_t)src & 15) != 0 || ((uintptr_t)kern_border & 15) != 0 || ((uintptr_t)kern_img & 15) != 0 ||
      ((uintptr_t)kc_img & 15) != 0 || ((uintptr_t)scratch & 15) != 0)
    return 1;
  uint8_t* sc = (uint8_t*)scratch;
  if (C % 256 == 0) {
    if (radius == 5) return launch_fused_shared<5, 2>(src, n_crops, h, w, C, kern_border, kern_img, kc_img, sg, ldk, dst, sc, st);
    return launch_fused_shared<3, 2>(src, n_crops, h, w, C, kern_border, kern_img, kc_img, sg, ldk, dst, sc, st);
  }
  if (radius == 5) return launch_fused_shared<5, 1>(src, n_crops, h, w, C, kern_border, kern_img, kc_img, sg, ldk, dst, sc, st);
  return launch_fused_shared<3, 1>(src, n_crops, h, w, C, kern_border, kern_img, kc_img, sg, ldk, dst, sc, st);
}
