// JBU adaptive convolution on tensor cores (bf16 path of cseg_jbu_apply).
//
//   out[y, x, c] = sum_{i,j} hr[reflect(y+i-R), reflect(x+j-R), c] * kern[(y,x)][i*D + j]        D = 2R+1
//
// (simfeatup_dev/upsamplers.py:269-274, semantics of adaptive_conv_py_simple :14-25.)  For a fixed
// output row y and tap row i this is a banded GEMM  out[16 px, C] += Wband[16 px, 32 pos] . hr[32 pos, C]
// with Wband[m][k] = kern[x0+m][i*D + (k-m)] for 0 <= k-m < D, which maps onto mma.sync m16n8k16
// (bf16 in, fp32 accumulate) at ~1/3 density -- still several times the CUDA-core FMA rate, and it takes
// the D*D multiply-adds per output off the critical path so the kernel can approach its HBM bound.
//
// CTA = RW output rows x 32 pixels x 128 channels, 4 warps = 2 pixel blocks x 2 channel halves
// (8 n-blocks of 8 channels each).  The RW+2R source rows stream through a cp.async ring in shared
// memory ([pos][channel], 16 B row padding => conflict-free ldmatrix.trans for the B fragments); every
// source row feeds up to RW output rows, so a B fragment is reused RW times.  The kernel weights of the
// tile are staged once in shared memory; A fragments are gathered from them.
#include "common.cuh"

namespace {

constexpr int RW = 4;      // output rows per CTA
constexpr int TX = 16;     // output pixels per CTA (one MMA pixel block; two CTAs are resident per SM)
constexpr int CS = 256;    // channels per CTA (4 warps x 64)
constexpr int NPOS = TX + 16;                 // source positions per row buffer (x0-R .. incl. K padding)
constexpr int PSTRIDE = CS * 2 + 16;          // bytes per position in the ring (padded)
constexpr int NST = 3;                        // ring stages
constexpr int ROW_BYTES = NPOS * PSTRIDE;
constexpr int ASTRIDE = 80;                   // bytes per row of a band tile (32 bf16 + pad: conflict-free ldmatrix)
constexpr int ATILE = 16 * ASTRIDE;           // one 16 x 32 band tile
constexpr int XB = TX / 16;                   // MMA pixel blocks per CTA
constexpr int NTHREADS = 128 * XB;

__device__ __forceinline__ int reflect1(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// CTA = RW output rows x TX pixels x 256 channels; warps = TX/16 pixel blocks x 4 channel quarters (8 n-blocks
// of 8 channels each).  The kernel weights of the tile are expanded ONCE into band tiles in shared memory
// (Wband[r][i][xb][m][k] = kern[px(m)][i*D + k - m] for 0 <= k-m < D, zero elsewhere), so an A fragment is
// two ldmatrix.x4 instead of a predicated gather, and it is shared by the four channel-quarter warps.
template <int R>
__global__ void __launch_bounds__(NTHREADS, 2 / XB) adaptive_conv_mma_kernel(const bf16* __restrict__ hr, int H2, int W2, int C,
                                                                        const bf16* __restrict__ kern, int ldk,
                                                                        bf16* __restrict__ dst) {
  pdl_grid_sync();
  constexpr int D = 2 * R + 1;
  constexpr int NSRC = RW + 2 * R;  // source rows per tile
  constexpr int WB_BYTES = RW * D * XB * ATILE;
  extern __shared__ __align__(16) uint8_t smem[];
  const uint32_t ring = (uint32_t)__cvta_generic_to_shared(smem);
  uint8_t* wb = smem + NST * ROW_BYTES;
  const uint32_t wband = ring + NST * ROW_BYTES;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int xb = warp % XB, cq = warp / XB;
  const int g = lane >> 2, tig = lane & 3;
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * RW;
  const int nslab = (C + CS - 1) / CS;
  const int crop = blockIdx.z / nslab, c0 = (blockIdx.z % nslab) * CS;
  const int cvalid = min(CS, C - c0);            // channels of this slab that exist
  const bool warp_on = cq * 64 < cvalid;         // warp-uniform: this channel quarter exists
  const bf16* hrc = hr + (size_t)crop * H2 * W2 * C + c0;

  // per-thread copy plan of a source row: the (position, channel chunk) pairs a thread moves are the same for
  // every row, so the reflect / index arithmetic is done once and a row costs LPT cp.async per thread
  constexpr int LPT = NPOS * (CS / 8) / NTHREADS;   // 6
  int src_off[LPT];                                  // element offset inside one source row, -1 = nothing to copy
  uint32_t dst_off[LPT];
#pragma unroll
  for (int k = 0; k < LPT; ++k) {
    const int e = tid + k * NTHREADS;
    const int p = e / (CS / 8), ch = (e % (CS / 8)) * 8;
    const int xx = reflect1(min(x0 - R + p, W2 - 1 + R), W2);
    src_off[k] = (ch < cvalid) ? xx * C + ch : -1;
    dst_off[k] = (uint32_t)(p * PSTRIDE + ch * 2);
  }
  auto load_row = [&](int sr) {
    const int yy = reflect1(min(y0 + sr - R, H2 - 1 + R), H2);
    const uint32_t base = ring + (sr % NST) * ROW_BYTES;
    const bf16* rowp = hrc + (size_t)yy * W2 * C;
#pragma unroll
    for (int k = 0; k < LPT; ++k)
      if (src_off[k] >= 0) cp_async16(base + dst_off[k], rowp + src_off[k]);
  };
#pragma unroll
  for (int s = 0; s < NST - 1; ++s) {
    if (s < NSRC) load_row(s);
    cp_async_commit();
  }
  // ---- expand the kernel weights of the tile into band tiles ----
  // the global loads are issued first so that their latency overlaps the zero fill
  constexpr int WPT = RW * TX * 16 / NTHREADS;       // 8 x 16-byte weight chunks per thread
  uint4 wv[WPT];
#pragma unroll
  for (int k = 0; k < WPT; ++k) {
    const int e = tid + k * NTHREADS;
    const int v = e % 16, px = (e / 16) % TX, r = e / (16 * TX);
    wv[k] = make_uint4(0, 0, 0, 0);
    if (y0 + r < H2 && x0 + px < W2 && v * 8 < ldk)
      wv[k] = __ldg(reinterpret_cast<const uint4*>(kern + (((size_t)crop * H2 + y0 + r) * W2 + x0 + px) * ldk + v * 8));
  }
  for (int e = tid; e < WB_BYTES / 16; e += NTHREADS) reinterpret_cast<uint4*>(wb)[e] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  {
    // NTHREADS % 16 == 0, so the 16-byte chunk index v = tid & 15 (hence the 8 taps a thread scatters) is the
    // same for all of its chunks: the tap -> (i, j) split is done once per thread
    const int v = tid & 15;
    int toff[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int t = v * 8 + q, i = t / D, j = t - i * D;
      toff[q] = (t < D * D) ? i * XB * ATILE + j * 2 : -1;
    }
#pragma unroll
    for (int k = 0; k < WPT; ++k) {
      const int e = tid + k * NTHREADS;
      const int px = (e / 16) % TX, r = e / (16 * TX);
      const unsigned short* hv = reinterpret_cast<const unsigned short*>(&wv[k]);
      const int m = px & 15, pxb = px >> 4;
      uint8_t* base = wb + (r * D * XB + pxb) * ATILE + m * (ASTRIDE + 2);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (toff[q] >= 0) *reinterpret_cast<unsigned short*>(base + toff[q]) = hv[q];
    }
  }

  float acc[RW][8][4];
#pragma unroll
  for (int r = 0; r < RW; ++r)
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[r][nb][e] = 0.f;

  const int lrow = lane & 7, lq = lane >> 3;
  const uint32_t b_lane_off = (uint32_t)(((lq & 1) * 8 + lrow) * PSTRIDE + (cq * 64 + (lq >> 1) * 8) * 2);
  // A fragment lanes: matrices (rows 0-7, k 0-7) (rows 8-15, k 0-7) (rows 0-7, k 8-15) (rows 8-15, k 8-15)
  const uint32_t a_lane_off = (uint32_t)(((lq & 1) * 8 + lrow) * ASTRIDE + (lq >> 1) * 16);

#pragma unroll 1
  for (int sr = 0; sr < NSRC; ++sr) {
    cp_async_wait<NST - 2>();
    __syncthreads();  // row sr has landed for everyone (and, at sr = 0, the band tiles are complete);
                      // slot (sr-1)%NST is free again
    if (sr + NST - 1 < NSRC) load_row(sr + NST - 1);
    cp_async_commit();
    if (!warp_on) continue;
    const uint32_t rowbase = ring + (sr % NST) * ROW_BYTES;
    if (sr >= RW - 1 && sr < D) {
      // interior source row: it feeds all RW output rows (tap rows i = sr - r).  Branch-free, so all sixteen
      // fragment loads are issued before the 64 MMAs and their latency overlaps the tensor pipe.
      uint32_t bfr[2][4][4], a[2][RW][4];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const uint32_t baddr = rowbase + (uint32_t)((xb * 16 + ks * 16) * PSTRIDE) + b_lane_off;
#pragma unroll
        for (int j2 = 0; j2 < 4; ++j2) ldsm_x4_trans(baddr + j2 * 32, bfr[ks][j2]);
#pragma unroll
        for (int r = 0; r < RW; ++r)
          ldsm_x4(wband + (uint32_t)(((r * D + (sr - r)) * XB + xb) * ATILE + ks * 32) + a_lane_off, a[ks][r]);
      }
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int r = 0; r < RW; ++r)
#pragma unroll
          for (int nb = 0; nb < 8; ++nb)
            mma_bf16(acc[r][nb], a[ks][r], bfr[ks][nb >> 1][(nb & 1) * 2], bfr[ks][nb >> 1][(nb & 1) * 2 + 1]);
    } else {
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        uint32_t bfr[4][4];
        const uint32_t baddr = rowbase + (uint32_t)((xb * 16 + ks * 16) * PSTRIDE) + b_lane_off;
#pragma unroll
        for (int j2 = 0; j2 < 4; ++j2) ldsm_x4_trans(baddr + j2 * 32, bfr[j2]);
#pragma unroll
        for (int r = 0; r < RW; ++r) {
          const int i = sr - r;  // tap row of output row r fed by this source row
          if (i < 0 || i >= D) continue;
          uint32_t a[4];
          ldsm_x4(wband + (uint32_t)(((r * D + i) * XB + xb) * ATILE + ks * 32) + a_lane_off, a);
#pragma unroll
          for (int nb = 0; nb < 8; ++nb) mma_bf16(acc[r][nb], a, bfr[nb >> 1][(nb & 1) * 2], bfr[nb >> 1][(nb & 1) * 2 + 1]);
        }
      }
    }
  }
  cp_async_wait<0>();
  if (!warp_on) return;
  // epilogue: c0,c1 -> (px g, ch 2tig,2tig+1); c2,c3 -> (px g+8, ...)
#pragma unroll
  for (int r = 0; r < RW; ++r) {
    const int y = y0 + r;
    if (y >= H2) continue;
#pragma unroll
    for (int hm = 0; hm < 2; ++hm) {
      const int x = x0 + xb * 16 + g + hm * 8;
      if (x >= W2) continue;
      bf16* o = dst + (((size_t)crop * H2 + y) * W2 + x) * C + c0 + cq * 64 + 2 * tig;
#pragma unroll
      for (int nb = 0; nb < 8; ++nb)
        if (cq * 64 + nb * 8 < cvalid)
          *reinterpret_cast<__nv_bfloat162*>(o + nb * 8) = __floats2bfloat162_rn(acc[r][nb][hm * 2], acc[r][nb][hm * 2 + 1]);
    }
  }
}

template <int R>
int launch_conv(const bf16* hr, int n_crops, int H2, int W2, int C, const bf16* kern, int ldk, bf16* dst, cudaStream_t st) {
  constexpr int D = 2 * R + 1;
  const int smem = NST * ROW_BYTES + RW * D * XB * ATILE;
  CSEG_SET_SMEM(adaptive_conv_mma_kernel<R>, smem);
  dim3 grid(cdiv(W2, TX), cdiv(H2, RW), n_crops * cdiv(C, CS));
  CSEG_REQUIRE(grid.z <= 65535, "jbu_apply(bf16): too many crop x channel slabs (%u)", grid.z);
  cseg_launch(adaptive_conv_mma_kernel<R>, dim3(grid), dim3(NTHREADS), smem, st, hr, H2, W2, C, kern, ldk, dst);
  CSEG_LAUNCH_CHECK("jbu_adaptive_conv_mma");
  return 0;
}

}  // namespace

int cseg_jbu_adaptive_conv_mma(const bf16* hr, int n_crops, int H2, int W2, int C, const bf16* kern, int ldk,
                               int radius, bf16* dst, cudaStream_t st) {
  CSEG_REQUIRE(C % 64 == 0, "jbu_apply(bf16): C=%d must be a multiple of 64", C);
  CSEG_REQUIRE(ldk % 8 == 0 && ldk <= 128, "jbu_apply(bf16): ldk=%d must be a multiple of 8 and <= 128", ldk);
  if (radius == 5) return launch_conv<5>(hr, n_crops, H2, W2, C, kern, ldk, dst, st);
  return launch_conv<3>(hr, n_crops, H2, W2, C, kern, ldk, dst, st);
}
