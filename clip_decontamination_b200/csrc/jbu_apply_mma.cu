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
constexpr int TX = 32;     // output pixels per CTA (2 MMA pixel blocks)
constexpr int CS = 128;    // channels per CTA
constexpr int NPOS = TX + 16;                 // source positions per row buffer (x0-R .. incl. K padding)
constexpr int PSTRIDE = CS * 2 + 16;          // bytes per position in the ring (padded)
constexpr int NST = 4;                        // ring stages
constexpr int WSTRIDE = 136;                  // bf16 elements per staged kernel row (128 + 8 pad)
constexpr int ROW_BYTES = NPOS * PSTRIDE;
constexpr int SMEM_BYTES = NST * ROW_BYTES + RW * TX * WSTRIDE * 2;

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

template <int R>
__global__ void __launch_bounds__(128, 2) adaptive_conv_mma_kernel(const bf16* __restrict__ hr, int H2, int W2, int C,
                                                                   const bf16* __restrict__ kern, int ldk,
                                                                   bf16* __restrict__ dst) {
  constexpr int D = 2 * R + 1;
  constexpr int NSRC = RW + 2 * R;  // source rows per tile
  extern __shared__ __align__(16) uint8_t smem[];
  const uint32_t ring = (uint32_t)__cvta_generic_to_shared(smem);
  bf16* Ws = reinterpret_cast<bf16*>(smem + NST * ROW_BYTES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int xb = warp & 1, chalf = warp >> 1;
  const int g = lane >> 2, tig = lane & 3;
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * RW;
  const int nslab = C / CS;
  const int crop = blockIdx.z / nslab, c0 = (blockIdx.z % nslab) * CS;
  const bf16* hrc = hr + (size_t)crop * H2 * W2 * C + c0;

  auto load_row = [&](int sr) {
    const int yy = reflect1(min(y0 + sr - R, H2 - 1 + R), H2);
    const uint32_t base = ring + (sr % NST) * ROW_BYTES;
    for (int e = tid; e < NPOS * (CS / 8); e += 128) {
      const int p = e / (CS / 8), ch = (e % (CS / 8)) * 8;
      const int xx = reflect1(min(x0 - R + p, W2 - 1 + R), W2);
      cp_async16(base + p * PSTRIDE + ch * 2, hrc + ((size_t)yy * W2 + xx) * C + ch);
    }
  };
#pragma unroll
  for (int s = 0; s < NST - 1; ++s) {
    if (s < NSRC) load_row(s);
    cp_async_commit();
  }
  // stage the kernel weights of the tile: RW x TX rows of ldk bf16 (zero rows outside the image)
  for (int e = tid; e < RW * TX * (128 / 8); e += 128) {
    const int v = e % 16, px = (e / 16) % TX, r = e / (16 * TX);
    uint4 val = make_uint4(0, 0, 0, 0);
    if (y0 + r < H2 && x0 + px < W2 && v * 8 < ldk)
      val = __ldg(reinterpret_cast<const uint4*>(kern + (((size_t)crop * H2 + y0 + r) * W2 + x0 + px) * ldk + v * 8));
    *reinterpret_cast<uint4*>(Ws + (r * TX + px) * WSTRIDE + v * 8) = val;
  }

  float acc[RW][8][4];
#pragma unroll
  for (int r = 0; r < RW; ++r)
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[r][nb][e] = 0.f;

  const int lrow = lane & 7, lq = lane >> 3;
  const uint32_t b_lane_off = (uint32_t)(((lq & 1) * 8 + lrow) * PSTRIDE + (chalf * 64 + (lq >> 1) * 8) * 2);
  const unsigned short* Wu = reinterpret_cast<const unsigned short*>(Ws);

#pragma unroll 1
  for (int sr = 0; sr < NSRC; ++sr) {
    cp_async_wait<NST - 2>();
    __syncthreads();  // row sr has landed for everyone; slot (sr-1)%NST is free again
    if (sr + NST - 1 < NSRC) load_row(sr + NST - 1);
    cp_async_commit();
    const uint32_t rowbase = ring + (sr % NST) * ROW_BYTES;
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
        // A fragment of the band: A[m][k] = w[px(m)][i*D + k + 16 ks - m]
        uint32_t a[4];
        const unsigned short* w0 = Wu + (r * TX + xb * 16 + g) * WSTRIDE + i * D;
        const unsigned short* w1 = w0 + 8 * WSTRIDE;
        const int jb = 2 * tig + 16 * ks;
        auto wv = [&](const unsigned short* w, int j) -> uint32_t { return (j >= 0 && j < D) ? (uint32_t)w[j] : 0u; };
        a[0] = wv(w0, jb - g) | (wv(w0, jb + 1 - g) << 16);
        a[1] = wv(w1, jb - g - 8) | (wv(w1, jb + 1 - g - 8) << 16);
        a[2] = wv(w0, jb + 8 - g) | (wv(w0, jb + 9 - g) << 16);
        a[3] = wv(w1, jb - g) | (wv(w1, jb + 1 - g) << 16);
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) mma_bf16(acc[r][nb], a, bfr[nb >> 1][(nb & 1) * 2], bfr[nb >> 1][(nb & 1) * 2 + 1]);
      }
    }
  }
  cp_async_wait<0>();
  // epilogue: c0,c1 -> (px g, ch 2tig,2tig+1); c2,c3 -> (px g+8, ...)
#pragma unroll
  for (int r = 0; r < RW; ++r) {
    const int y = y0 + r;
    if (y >= H2) continue;
#pragma unroll
    for (int hm = 0; hm < 2; ++hm) {
      const int x = x0 + xb * 16 + g + hm * 8;
      if (x >= W2) continue;
      bf16* o = dst + (((size_t)crop * H2 + y) * W2 + x) * C + c0 + chalf * 64 + 2 * tig;
#pragma unroll
      for (int nb = 0; nb < 8; ++nb)
        *reinterpret_cast<__nv_bfloat162*>(o + nb * 8) = __floats2bfloat162_rn(acc[r][nb][hm * 2], acc[r][nb][hm * 2 + 1]);
    }
  }
}

}  // namespace

int cseg_jbu_adaptive_conv_mma(const bf16* hr, int n_crops, int H2, int W2, int C, const bf16* kern, int ldk,
                               int radius, bf16* dst, cudaStream_t st) {
  CSEG_REQUIRE(C % CS == 0, "jbu_apply(bf16): C=%d must be a multiple of %d", C, CS);
  CSEG_REQUIRE(ldk % 8 == 0 && ldk <= 128, "jbu_apply(bf16): ldk=%d must be a multiple of 8 and <= 128", ldk);
  dim3 grid(cdiv(W2, TX), cdiv(H2, RW), n_crops * (C / CS));
  CSEG_REQUIRE(grid.z <= 65535, "jbu_apply(bf16): too many crop x channel slabs (%u)", grid.z);
  if (radius == 5) {
    CSEG_SET_SMEM(adaptive_conv_mma_kernel<5>, SMEM_BYTES);
    adaptive_conv_mma_kernel<5><<<grid, 128, SMEM_BYTES, st>>>(hr, H2, W2, C, kern, ldk, dst);
  } else {
    CSEG_SET_SMEM(adaptive_conv_mma_kernel<3>, SMEM_BYTES);
    adaptive_conv_mma_kernel<3><<<grid, 128, SMEM_BYTES, st>>>(hr, H2, W2, C, kern, ldk, dst);
  }
  CSEG_LAUNCH_CHECK("jbu_adaptive_conv_mma");
  return 0;
}
