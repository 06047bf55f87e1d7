// JBU range kernel on tensor cores (bf16 path of cseg_jbu_range_kernel).
//
//   logit[p][t=(i,j)] = pos_temp * <proj(reflect(p + t - R)), proj(p)>            (upsamplers.py:230-238)
//   kern[p][t] = softmax_t(logit) * gauss(t) / max(sum_t softmax*gauss, 1e-7)    (:240-251,258-262)
//
// For 16 query pixels of one image row and one tap row i, the 16 x (16+2R) key.query products are a small
// GEMM: S[16 queries, 8-position blocks] = Q[16, 32] . K[positions, 32]^T  -> mma.sync m16n8k16 with fp16
// operands (the reference computes these projections in fp16 under autocast, segmentor.py:370) and fp32
// accumulation.  Only the band 0 <= pos - query < D of each block is used.  One pass over the D tap rows with a
// running maximum (online softmax): exp, both sums, and exp * gauss staged per warp in shared memory as fp16
// (values <= 1, relative to the running maximum of their tap row).  The normalisation -- including the factor
// 2^(m_row - m_final) -- is applied while the staged rows are copied out as full 16-byte vectors of the
// [pixels, ldk] kernel matrix (taps, then the 3 guidance channels, then zeros).
#include "common.cuh"
#include "jbu_share.cuh"
#include <cuda_fp16.h>

namespace {

constexpr int TXR = 32, TYR = 8;   // CTA tile: 32 x 8 query pixels, one warp per row
constexpr int KD = 32;             // projection width
constexpr int PROW = KD * 2 + 16;  // bytes per staged position (padded: conflict-free ldmatrix)
constexpr int SPAD = 8;            // staging row padding (halves): rows 4 banks apart, 16-byte aligned
constexpr int MROWS = 12;          // per-pixel running maxima of the tap rows (D <= 11, padded)

__device__ __forceinline__ int reflect1(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// BORDER = false: every pixel of n regions of gh x gw (one crop each, or the whole canvas as one region), dense output
// [region][gh][gw].  BORDER = true (jbu_share.cuh): only the border frame of every crop -- top / bottom `fb` rows and
// the left / right 16 columns -- read from the IMAGE-LEVEL projection / guidance buffers at the crop's origin (reflect
// padding is relative to the crop), compact output.
template <int R, int LDK, bool BORDER>
__global__ void __launch_bounds__(TYR * 32) range_kernel_mma(const __half* __restrict__ proj,
                                                             const float4* __restrict__ guid, int gh, int gw,
                                                             float pos_temp, float inv2s2, bf16* __restrict__ kern,
                                                             int ldk, const ShareGeom geom, int frame) {
  pdl_grid_sync();
  constexpr int D = 2 * R + 1, D2 = D * D;
  constexpr int NB = (16 + 2 * R + 7) / 8;       // 8-position blocks per 16-query block
  constexpr int HR = TYR + 2 * R;                // halo rows
  constexpr int NPOS = 16 + NB * 8;              // staged positions per halo row
  extern __shared__ __align__(16) uint8_t rsm[];
  const uint32_t psm = (uint32_t)__cvta_generic_to_shared(rsm);
  float* gauss = reinterpret_cast<float*>(rsm + HR * NPOS * PROW);
  bf16* stage = reinterpret_cast<bf16*>(rsm + HR * NPOS * PROW + ((D2 * 4 + 15) & ~15));
  float* mrow = reinterpret_cast<float*>(rsm + HR * NPOS * PROW + ((D2 * 4 + 15) & ~15) + TYR * 16 * (LDK + SPAD) * 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;
  const int crop = blockIdx.z;
  int y0, x0, nxb = TXR / 16, ylim = gh, pitch = gw;
  const __half* pc;
  const float4* gc;
  if (!BORDER) {
    y0 = blockIdx.y * TYR;
    x0 = blockIdx.x * TXR;
    pc = proj + (size_t)crop * gh * gw * KD;
    gc = guid + (size_t)crop * gh * gw;
  } else {
    const int ntx = (gw + TXR - 1) / TXR, t1 = (frame / TYR) * ntx, idx = blockIdx.x;
    if (idx < t1) {                                            // top strip
      y0 = (idx / ntx) * TYR;
      x0 = (idx % ntx) * TXR;
    } else if (idx < 2 * t1) {                                 // bottom strip
      y0 = gh - frame + ((idx - t1) / ntx) * TYR;
      x0 = ((idx - t1) % ntx) * TXR;
    } else {                                                   // left / right 16 columns of the rows in between
      const int j = idx - 2 * t1;
      y0 = frame + (j >> 1) * TYR;
      x0 = (j & 1) ? gw - 16 : 0;
      nxb = 1;
      ylim = gh - frame;
    }
    pitch = geom.pitch;
    const size_t org = (size_t)(geom.wins[crop * 4] >> geom.shift) * pitch + (geom.wins[crop * 4 + 1] >> geom.shift);
    pc = proj + org * KD;
    gc = guid + org;
  }

  for (int e = tid; e < HR * NPOS * 4; e += TYR * 32) {      // 4 x 16 B per position, all in flight (cp.async)
    const int ch = e & 3, pos = (e >> 2) % NPOS, hy = (e >> 2) / NPOS;
    const int yy = reflect1(min(y0 - R + hy, gh - 1 + R), gh), xx = reflect1(min(x0 - R + pos, gw - 1 + R), gw);
    const uint32_t dst = psm + (uint32_t)((hy * NPOS + pos) * PROW + ch * 16);
    const __half* src = pc + ((size_t)yy * pitch + xx) * KD + ch * 8;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  // get_spatial_kernel: exp(-(dy^2 + dx^2) / 2 sigma^2) on linspace(-1, 1, D)^2 is separable: gauss[i] holds the
  // 1-D factor; the column factors of a thread's fragment elements live in registers
  for (int t = tid; t < D; t += TYR * 32) {
    const float d1 = -1.f + 2.f * t / (D - 1);
    gauss[t] = __expf(-(d1 * d1) * inv2s2);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const int y = y0 + warp;
  bf16* st = stage + warp * 16 * (LDK + SPAD);
  const int q = lane >> 3, rr = lane & 7;
  // the staging columns behind the D2 taps are never written by the tap loop: zero them once, so that the copy-out can
  // weight them with 0 instead of selecting (uninitialised shared memory may hold NaN patterns)
  if (lane < 16) {
    __half* zr = reinterpret_cast<__half*>(st) + lane * (LDK + SPAD);
    for (int c = D2; c < LDK + SPAD; ++c) zr[c] = __float2half(0.f);
  }
  __syncwarp();
#pragma unroll 1
  for (int xb = 0; xb < TXR / 16; ++xb) {
    const int xq0 = x0 + xb * 16;
    if (y >= ylim || xq0 >= gw || xb >= nxb) break;          // warp-uniform
    uint32_t a[2][4];
    {
      const uint32_t base = psm + (uint32_t)(((warp + R) * NPOS + xb * 16 + R + (q & 1) * 8 + rr) * PROW + (q >> 1) * 16);
      ldsm4(base, a[0]);
      ldsm4(base + 32, a[1]);
    }
    float se[2] = {0.f, 0.f}, sg[2] = {0.f, 0.f}, inv[2] = {0.f, 0.f};
    // tap column j of fragment element (nb, e) does not depend on the tap row i: hoist it (and its validity)
    int jj[NB][4];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = nb * 8 + 2 * tig + (e & 1) - (g + (e >> 1) * 8);
        jj[nb][e] = (j >= 0 && j < D) ? j : -1;
      }
    float gx[NB][4];                                           // column factor of the spatial Gaussian per element
#pragma unroll
    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
      for (int e = 0; e < 4; ++e) gx[nb][e] = jj[nb][e] >= 0 ? gauss[jj[nb][e]] : 0.f;
    const uint32_t rowb0 = psm + (uint32_t)((warp * NPOS + xb * 16 + rr) * PROW + q * 16);
    // ---- single pass over the tap rows with a running maximum (online softmax): exp, both sums, and the unnormalised
    //      exp * gauss_x (fp16, <= 1) into the staging rows.  Every tap row i is staged relative to the running maximum at
    //      that time, which is kept per (pixel, i); that factor, the row factor gauss_y[i] of the separable Gaussian and the
    //      normalisation are folded in when the rows are copied out. ----
    // exp(pos_temp * s - m) = 2^(s * pt2 - m'): ONE FFMA + MUFU.EX2 per tap, the scale is never applied to S itself.  The
    // running maximum needs a positive scale: for a negative pos_temp the query fragments change sign (exact in fp16).
    float pt2 = pos_temp * 1.4426950408889634f;
    if (pt2 < 0.f) {                                           // uniform
      pt2 = -pt2;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int r4 = 0; r4 < 4; ++r4) a[ks][r4] ^= 0x80008000u;
    }
    float mrun[2] = {-INFINITY, -INFINITY};                    // running maxima of rows g, g+8 in log2 units
    __half* sth = reinterpret_cast<__half*>(st);
    float* mrw = mrow + warp * 16 * MROWS;
#pragma unroll 1
    for (int i = 0; i < D; ++i) {
      const uint32_t rowb = rowb0 + (uint32_t)(i * NPOS * PROW);
      __half* sti = sth + i * D;
      float S[NB][4];
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        uint32_t b[4];
        ldsm4(rowb + nb * 8 * PROW, b);
        S[nb][0] = S[nb][1] = S[nb][2] = S[nb][3] = 0.f;
        mma_f16(S[nb], a[0], b[0], b[1]);
        mma_f16(S[nb], a[1], b[2], b[3]);
      }
      // band mask once per element: everything after it (maximum, exponent, sums) is unpredicated; 2^(-inf) = 0
      float mi[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nb = 0; nb < NB; ++nb)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if ((e < 2 && nb * 8 >= 8 + D - 1) || (e >= 2 && nb * 8 + 8 <= 8)) continue;     // compile-time: never in the band
          S[nb][e] = jj[nb][e] >= 0 ? S[nb][e] : -INFINITY;
          mi[e >> 1] = fmaxf(mi[e >> 1], S[nb][e]);
        }
      float nm[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        mi[h] = fmaxf(mi[h], __shfl_xor_sync(0xffffffffu, mi[h], 1));
        mi[h] = fmaxf(mi[h], __shfl_xor_sync(0xffffffffu, mi[h], 2));
        const float mnew = fmaxf(mrun[h], __fmul_rn(mi[h], pt2));
        float corr;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(corr) : "f"(mrun[h] - mnew));   // 0 on the first row (mrun = -inf)
        se[h] = __fmul_rn(se[h], corr);
        sg[h] = __fmul_rn(sg[h], corr);
        mrun[h] = mnew;
        nm[h] = -mnew;
        if (tig == 0) mrw[(g + h * 8) * MROWS + i] = mnew;
      }
      // Rows g (e < 2) only reach positions [g, g + D) of the 8 NB-blocks, rows g + 8 (e >= 2) positions [g + 8, g + 8 + D):
      // the last block can never hold a tap of the top half, the first block never one of the bottom half.
      float sgr[2] = {0.f, 0.f};
#pragma unroll
      for (int nb = 0; nb < NB; ++nb)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          if ((hh == 0 && nb * 8 >= 8 + D - 1) || (hh == 1 && nb * 8 + 8 <= 8)) continue;   // compile-time
          const int e0 = hh * 2;
          float ex0, ex1;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex0) : "f"(fmaf(S[nb][e0], pt2, nm[hh])));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex1) : "f"(fmaf(S[nb][e0 + 1], pt2, nm[hh])));
          // explicit roundings: the BORDER and dense instantiations must not differ in FMA contraction (their results are
          // required to be bit-identical, tests/test_engine_gpu.py::test_jbu_shared_kernels_equal_per_crop)
          const float w0 = __fmul_rn(ex0, gx[nb][e0]), w1 = __fmul_rn(ex1, gx[nb][e0 + 1]);
          se[hh] = __fadd_rn(se[hh], __fadd_rn(ex0, ex1));
          sgr[hh] = __fadd_rn(sgr[hh], __fadd_rn(w0, w1));
          const __half2 wp = __floats2half2_rn(w0, w1);                  // one F2FP for the pair
          __half* srow = sti + (g + hh * 8) * (LDK + SPAD);
          if (jj[nb][e0] >= 0) srow[jj[nb][e0]] = __low2half(wp);
          if (jj[nb][e0 + 1] >= 0) srow[jj[nb][e0 + 1]] = __high2half(wp);
        }
      const float gy = gauss[i];
      sg[0] = fmaf(gy, sgr[0], sg[0]);
      sg[1] = fmaf(gy, sgr[1], sg[1]);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      se[h] += __shfl_xor_sync(0xffffffffu, se[h], 1);
      se[h] += __shfl_xor_sync(0xffffffffu, se[h], 2);
      sg[h] += __shfl_xor_sync(0xffffffffu, sg[h], 1);
      sg[h] += __shfl_xor_sync(0xffffffffu, sg[h], 2);
      const float ise = 1.0f / se[h];
      inv[h] = ise / fmaxf(__fmul_rn(sg[h], ise), 1e-7f);               // softmax, then / sum(softmax*gauss).clamp(1e-7)
    }
    __syncwarp();
    // ---- normalise while copying out: taps * inv[row] * gauss_y[i] * 2^(m_i - m_final), then the 3 guidance channels, then
    //      zeros; coalesced 16-byte row stores of the [pixels, ldk] kernel matrix.  A lane keeps its 16-byte column chunk v for
    //      all iterations (32 lanes = 32 / CH pixels x CH chunks), so everything that depends on v only is hoisted. ----
    {
      constexpr int CH = LDK / 8;                              // 16-byte chunks per pixel row
      static_assert(CH == 16 || CH == 8, "copy-out assumes 8 or 16 chunks of 8 taps per pixel");
      const int v = lane % CH;
      // the 8 taps of chunk v lie in tap rows ia and (from element bnd on) ia + 1
      const int ia = (v * 8) / D, bnd = (ia + 1) * D - v * 8;
      const int ra = min(ia, D - 1), rb = min(ia + 1, D - 1);
      const float ga = gauss[ra], gb = gauss[rb];
      float wa[8], wb[8];                                      // 0 / 1 selectors: element k takes factor a, b or is padding
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const bool ok = v * 8 + k < D2;
        wa[k] = (ok && k < bnd) ? 1.f : 0.f;
        wb[k] = (ok && k >= bnd) ? 1.f : 0.f;
      }
      const bool has_guid = v == D2 / 8;
#pragma unroll
      for (int it = 0; it < CH / 2; ++it) {
        const int px = lane / CH + (32 / CH) * it;             // px < 8 <=> it < CH / 4 (compile time)
        const int src = (px & 7) * 4;
        const float sc = __shfl_sync(0xffffffffu, it < CH / 4 ? inv[0] : inv[1], src);
        const float mf = __shfl_sync(0xffffffffu, it < CH / 4 ? mrun[0] : mrun[1], src);
        float fa, fb;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(fa) : "f"(mrw[px * MROWS + ra] - mf));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(fb) : "f"(mrw[px * MROWS + rb] - mf));
        fa = __fmul_rn(fa, __fmul_rn(sc, ga));
        fb = __fmul_rn(fb, __fmul_rn(sc, gb));
        const uint4 raw = *reinterpret_cast<const uint4*>(sth + px * (LDK + SPAD) + v * 8);
        const __half2* hp = reinterpret_cast<const __half2*>(&raw);
        float f[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 t2 = __half22float2(hp[k]);
          f[2 * k] = __fmul_rn(t2.x, fmaf(wa[2 * k], fa, __fmul_rn(wb[2 * k], fb)));
          f[2 * k + 1] = __fmul_rn(t2.y, fmaf(wa[2 * k + 1], fa, __fmul_rn(wb[2 * k + 1], fb)));
        }
        if (has_guid) {                                        // columns D2 .. D2+2: guidance (RGB) of this pixel
          const int x = min(xq0 + px, gw - 1);
          const float4 gv = gc[(size_t)y * pitch + x];
          f[D2 % 8] = gv.x;
          f[D2 % 8 + 1] = gv.y;
          f[D2 % 8 + 2] = gv.z;
        }
        if (xq0 + px < gw) {
          uint4 o;
          __nv_bfloat162* op = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
          for (int k = 0; k < 4; ++k) op[k] = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
          const size_t orow = BORDER ? (size_t)crop * border_rows(gh, gw, frame) + border_index(y, xq0 + px, gh, gw, frame)
                                     : ((size_t)crop * gh + y) * gw + xq0 + px;
          *reinterpret_cast<uint4*>(kern + orow * ldk + v * 8) = o;
        }
      }
    }
    __syncwarp();
  }
}

// range_proj (upsamplers.py:209-214) with fp16 output for the tensor-core range kernel:
//   proj[p] = W3 . gelu(W0 . rgb[p] + b0) + b3        (two 1x1 convs, 3 -> 32 -> 32)
// One warp per 16 pixels.  The hidden layer is evaluated directly in the mma.sync A-fragment layout (16 GELUs per
// thread, erf form to 3e-7), the 32x32 second layer runs on the tensor cores with both operands split into
// fp16 hi + lo parts (hi.hi + hi.lo + lo.hi: products carry ~21 bits, accumulation is fp32), so the result equals
// the fp32 evaluation to ~1e-6 before the final fp16 rounding.
__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
  __half2 h = __halves2half2(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// guidance (adaptive_avg_pool2d of every crop to gh x gw, upsamplers.py:316) fused with the projection: the pooled RGB is
// computed by the lane quad of each pixel (lane tig = channel), written out as the guidance tensor (the range kernel and
// the kernel fix-up read it) and fed straight into the MLP.  FUSED = false is the stand-alone projection.
template <bool FUSED>
__global__ void __launch_bounds__(256) range_proj_f16_kernel(const float4* __restrict__ guid, float4* __restrict__ guid_out,
                                                             const ImgView img,
                                                             const int32_t* __restrict__ wins, int crop_h, int crop_w,
                                                             int pad_top, int pad_left, int gh, int gw, int n_pix,
                                                             const float* __restrict__ w0, const float* __restrict__ b0,
                                                             const float* __restrict__ w3, const float* __restrict__ b3,
                                                             __half* __restrict__ proj) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  // per-thread constants: the 8 hidden units (k = ks*16 + 2 tig + {0, 1, 8, 9}) of the A fragments ...
  float wa[2][4][4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = ks * 16 + 2 * tig + (q & 1) + (q >> 1) * 8;
      wa[ks][q][0] = w0[k * 3];
      wa[ks][q][1] = w0[k * 3 + 1];
      wa[ks][q][2] = w0[k * 3 + 2];
      wa[ks][q][3] = b0[k];
    }
  // ... and the B fragments of W3 (B[k][n] = w3[n][k]), hi and lo halves
  uint32_t bh[4][2][2], bl[4][2][2];
  float bias[4][2];
#pragma unroll
  for (int nb = 0; nb < 4; ++nb) {
    const int n = nb * 8 + g;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int k = ks * 16 + 2 * tig + r * 8;
        const float v0 = w3[n * KD + k], v1 = w3[n * KD + k + 1];
        const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
        bh[nb][ks][r] = pack_h2(h0, h1);
        bl[nb][ks][r] = pack_h2(__float2half_rn(v0 - __half2float(h0)), __float2half_rn(v1 - __half2float(h1)));
      }
    bias[nb][0] = b3[nb * 8 + 2 * tig];
    bias[nb][1] = b3[nb * 8 + 2 * tig + 1];
  }
  const int n_tiles = (n_pix + 15) / 16;
  for (int tile = warp_global; tile < n_tiles; tile += n_warps) {
    const int p0 = tile * 16 + g, p1 = p0 + 8;
    float4 g0, g1;
    if (!FUSED) {
      g0 = guid[min(p0, n_pix - 1)];
      g1 = guid[min(p1, n_pix - 1)];
    } else {
      // pixel coordinates: one division per tile (warp-uniform); the 16 pixels of a tile wrap at most once (gw >= 16)
      const int base = tile * 16, gx0 = base % gw, t0 = base / gw, gy0 = t0 % gh, crop0 = t0 / gh;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        int gx = gx0 + g + hh * 8, gy = gy0, crop = crop0;
        if (gx >= gw) {
          gx -= gw;
          if (++gy == gh) { gy = 0; ++crop; }
        }
        if (tile * 16 + g + hh * 8 >= n_pix) { gx = 0; gy = 0; crop = 0; }   // padding pixels of the last tile: any valid cell
        const int y1 = wins[crop * 4], x1 = wins[crop * 4 + 1], wh = wins[crop * 4 + 2], ww = wins[crop * 4 + 3];
        // adaptive pooling window: [floor(i*in/out), ceil((i+1)*in/out))
        const int ys = (gy * crop_h) / gh, ye = ((gy + 1) * crop_h + gh - 1) / gh;
        const int xs = (gx * crop_w) / gw, xe = ((gx + 1) * crop_w + gw - 1) / gw;
        // the four lanes of the quad split the window rows; all three channels per lane, then a quad reduction
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        for (int y = ys + tig; y < ye; y += 4) {
          const int cy = y - pad_top;
          if (cy < 0 || cy >= wh) continue;
          const long long r0 = img_row_off(img, y1 + cy);
          for (int x = xs; x < xe; ++x) {
            const int cx = x - pad_left;
            if (cx >= 0 && cx < ww) {
              a0 += img_at(img, 0, r0, x1 + cx);
              a1 += img_at(img, 1, r0, x1 + cx);
              a2 += img_at(img, 2, r0, x1 + cx);
            }
          }
        }
        a0 += __shfl_xor_sync(0xffffffffu, a0, 1); a0 += __shfl_xor_sync(0xffffffffu, a0, 2);
        a1 += __shfl_xor_sync(0xffffffffu, a1, 1); a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
        a2 += __shfl_xor_sync(0xffffffffu, a2, 1); a2 += __shfl_xor_sync(0xffffffffu, a2, 2);
        const float inv = 1.0f / (float)((ye - ys) * (xe - xs));
        if (hh == 0) g0 = make_float4(a0 * inv, a1 * inv, a2 * inv, 0.f);
        else g1 = make_float4(a0 * inv, a1 * inv, a2 * inv, 0.f);
      }
      if (tig == 0) {
        if (p0 < n_pix) guid_out[p0] = g0;
        if (p1 < n_pix) guid_out[p1] = g1;
      }
    }
    uint32_t ah[2][4], al[2][4];                       // A fragments: (row g | g+8) x (k pair | k pair + 8)
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int r = 0; r < 2; ++r) {                    // r: k pair (0,1) or (8,9)
        float h[2][2];                                 // [row][element of the pair]
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float* w = wa[ks][r * 2 + e];
          const float z0 = fmaf(w[0], g0.x, fmaf(w[1], g0.y, fmaf(w[2], g0.z, w[3])));
          const float z1 = fmaf(w[0], g1.x, fmaf(w[1], g1.y, fmaf(w[2], g1.z, w[3])));
          h[0][e] = FUSED ? gelu_tanh(z0) : gelu_fast(z0);   // tanh form (|err| <= 4.8e-4, the size of the fp16 rounding the
          h[1][e] = FUSED ? gelu_tanh(z1) : gelu_fast(z1);   // reference's autocast applies to these activations)
        }
#pragma unroll
        for (int row = 0; row < 2; ++row) {
          const __half x0 = __float2half_rn(h[row][0]), x1 = __float2half_rn(h[row][1]);
          ah[ks][r * 2 + row] = pack_h2(x0, x1);
          al[ks][r * 2 + row] = pack_h2(__float2half_rn(h[row][0] - __half2float(x0)),
                                        __float2half_rn(h[row][1] - __half2float(x1)));
        }
      }
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
      float d[4] = {bias[nb][0], bias[nb][1], bias[nb][0], bias[nb][1]};
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        mma_f16(d, al[ks], bh[nb][ks][0], bh[nb][ks][1]);
        mma_f16(d, ah[ks], bl[nb][ks][0], bl[nb][ks][1]);
        mma_f16(d, ah[ks], bh[nb][ks][0], bh[nb][ks][1]);
      }
      if (p0 < n_pix) *reinterpret_cast<__half2*>(proj + (size_t)p0 * KD + nb * 8 + 2 * tig) = __floats2half2_rn(d[0], d[1]);
      if (p1 < n_pix) *reinterpret_cast<__half2*>(proj + (size_t)p1 * KD + nb * 8 + 2 * tig) = __floats2half2_rn(d[2], d[3]);
    }
  }
}

template <int R, int LDK>
int launch(const __half* proj, const float* guid, int n_crops, int gh, int gw, float pos_temp, float inv2s2, bf16* kern,
           int ldk, cudaStream_t st, const ShareGeom* sg = nullptr, int fb = 0) {
  constexpr int D2 = (2 * R + 1) * (2 * R + 1), NB = (16 + 2 * R + 7) / 8, HR = TYR + 2 * R, NPOS = 16 + NB * 8;
  const int smem = HR * NPOS * PROW + ((D2 * 4 + 15) & ~15) + TYR * 16 * (LDK + SPAD) * 2 + TYR * 16 * MROWS * 4;
  if (sg == nullptr) {
    CSEG_SET_SMEM((range_kernel_mma<R, LDK, false>), smem);
    dim3 grid(cdiv(gw, TXR), cdiv(gh, TYR), n_crops);
    ShareGeom none = {nullptr, 0, 0};
    cseg_launch(range_kernel_mma<R, LDK, false>, dim3(grid), dim3(TYR * 32), smem, st, proj, (const float4*)guid, gh, gw, pos_temp,
                inv2s2, kern, ldk, none, 0);
  } else {
    CSEG_SET_SMEM((range_kernel_mma<R, LDK, true>), smem);
    const int tiles = 2 * (fb / TYR) * cdiv(gw, TXR) + 2 * cdiv(gh - 2 * fb, TYR);
    dim3 grid(tiles, 1, n_crops);
    cseg_launch(range_kernel_mma<R, LDK, true>, dim3(grid), dim3(TYR * 32), smem, st, proj, (const float4*)guid, gh, gw, pos_temp,
                inv2s2, kern, ldk, *sg, fb);
  }
  CSEG_LAUNCH_CHECK("jbu_range_kernel_mma");
  return 0;
}

}  // namespace

int cseg_jbu_range_proj_f16(const float* guid, int n_pix, const float* w0, const float* b0, const float* w3,
                            const float* b3, void* proj, cudaStream_t st) {
  const int blocks = (int)std::min<long long>(cdiv(cdiv(n_pix, 16), 8), (long long)sm_count() * 8);
  ImgView none = {};
  cseg_launch(range_proj_f16_kernel<false>, dim3(blocks), dim3(256), 0, st, (const float4*)guid, (float4*)nullptr,
              none, (const int32_t*)nullptr, 0, 0, 0, 0, 1, 1, n_pix, w0, b0, w3, b3, (__half*)proj);
  CSEG_LAUNCH_CHECK("jbu_range_proj_f16");
  return 0;
}

int cseg_jbu_guidance_proj_f16(const ImgView& img, const int32_t* windows, int n_crops, int crop_h, int crop_w,
                               int pad_top, int pad_left, int gh, int gw, const float* w0, const float* b0, const float* w3,
                               const float* b3, float* guid, void* proj, cudaStream_t st) {
  const int n_pix = n_crops * gh * gw;
  const int blocks = (int)std::min<long long>(cdiv(cdiv(n_pix, 16), 8), (long long)sm_count() * 8);
  cseg_launch(range_proj_f16_kernel<true>, dim3(blocks), dim3(256), 0, st, (const float4*)nullptr, (float4*)guid, img,
              windows, crop_h, crop_w, pad_top, pad_left, gh, gw, n_pix, w0, b0, w3, b3, (__half*)proj);
  CSEG_LAUNCH_CHECK("jbu_guidance_proj_f16");
  return 0;
}

// returns 1 when (radius, ldk) is not covered.  sg != nullptr: border frames only (jbu_share.cuh), fb % 8 == 0.
int cseg_jbu_range_kernel_mma(const void* proj_f16, const float* guid, int n_crops, int gh, int gw, int radius,
                              float pos_temp, float inv2s2, void* kern, int kwidth, int ldk, cudaStream_t st,
                              const ShareGeom* sg, int fb) {
  if (ldk % 8 != 0 || ((uintptr_t)kern & 15) != 0) return 1;
  if (sg != nullptr && (fb <= 0 || fb % TYR != 0 || gh <= 2 * fb || gw < 32)) return 1;
  if (radius == 5 && kwidth == 128)
    return launch<5, 128>((const __half*)proj_f16, guid, n_crops, gh, gw, pos_temp, inv2s2, (bf16*)kern, ldk, st, sg, fb);
  if (radius == 3 && kwidth == 64)
    return launch<3, 64>((const __half*)proj_f16, guid, n_crops, gh, gw, pos_temp, inv2s2, (bf16*)kern, ldk, st, sg, fb);
  return 1;
}
