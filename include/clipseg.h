/*
 * clipseg.h -- C ABI of libclipseg.so, the B200 (sm_100a) kernels behind the dense CLIP
 * segmentation hot path of UserNameUnavailableIsUnavailable/CLIP-Decontamination.
 *
 * The reference is pure Python and has no FFI; each entry point below names the reference
 * function (file:line, relative to the reference root) whose arithmetic it replaces.  The
 * reference-side binding is a ctypes stub; the package's own binding is
 * clip_decontamination_b200/_lib.py.
 *
 * Conventions
 *   - every function returns 0 on success, a negative CSEG_E* code on failure; the message is
 *     available through cseg_last_error() (thread-local).  No exception crosses the boundary.
 *   - all pointers are DEVICE pointers unless the name ends in _host; nothing is allocated
 *     inside; every launcher takes the CUDA stream (a cudaStream_t passed as void*).
 *   - dtype arguments are CSEG_F32 / CSEG_BF16; "T" in a comment means that dtype.
 *   - matrices are row-major; "ld*" are leading dimensions in ELEMENTS.
 *   - token tensors are crop-major: row = crop * L + token.
 */
#ifndef CLIPSEG_H_
#define CLIPSEG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSEG_VERSION 200

#if defined(__GNUC__)
#define CSEG_API __attribute__((visibility("default")))
#else
#define CSEG_API
#endif

enum { CSEG_F32 = 0, CSEG_BF16 = 1, CSEG_F16 = 2 /* JBU range projections only */, CSEG_U8 = 3 /* cseg_image only */ };
enum { CSEG_OK = 0, CSEG_EINVAL = -1, CSEG_ECUDA = -2, CSEG_EUNSUPPORTED = -3 };
/* epilogue activations of cseg_gemm */
enum { CSEG_ACT_NONE = 0, CSEG_ACT_GELU = 1, CSEG_ACT_QUICKGELU = 2 };
/* attention modes of cseg_attention: open_clip/transformer.py:858-908 */
enum {
  CSEG_ATTN_STD = 0,          /* softmax(q k^T * s) v           nn.MultiheadAttention, :204,218-232 */
  CSEG_ATTN_EXPERIMENTAL = 1, /* softmax(softmax((kk+qq) s) + M) v                        :896-902 */
  CSEG_ATTN_SCLIP = 2,        /* softmax(qq s + M) + softmax(kk s + M)                    :870-877 */
  CSEG_ATTN_CLEARCLIP = 3,    /* softmax(qq s + M)                                        :903-908 */
  CSEG_ATTN_SFP = 4,          /* softmax(0.5 (qq+kk) s + M)                               :888-895 */
  CSEG_ATTN_VANILLA = 5,      /* softmax(qk s + M)                                        :858-863 */
  CSEG_ATTN_SEGEARTH = 6,     /* SCLIP + softmax(vv s + M)                                :878-887 */
  CSEG_ATTN_MASKCLIP = 7,     /* identity attention                                       :864-869 */
  CSEG_ATTN_CAUSAL = 8        /* softmax(q k^T * s + causal mask) v: text tower, transformer.py:1047-1053, model.py:295 */
};

CSEG_API int cseg_version(void);
/* copies the calling thread's last error message into buf (NUL terminated); returns its length */
CSEG_API int cseg_last_error(char* buf, size_t n);
/* number of kernel launches issued through this library by the calling process so far */
CSEG_API long long cseg_launch_count(void);

/* ---- input side (N1): mmseg SegDataPreProcessor arithmetic, segmentor.py:64-67 -------------
 * Image descriptor read by the three kernels that touch the input (cseg_patchify, cseg_jbu_guidance,
 * cseg_jbu_guidance_proj).  A HOST struct passed by pointer; `data` is a DEVICE pointer.  The canvas has H rows
 * and W columns and is a vertical stack of H / img_h equally sized images (img_h == H: one image); canvas row Y is
 * row Y % img_h of image Y / img_h.  Element (c, Y, x) of the normalised RGB image the reference feeds to
 * predict() (segmentor.py:453-468) lives at
 *     data[(Y / img_h) * stride_img + chan[c] * stride_c + (Y % img_h) * stride_y + x * stride_x]
 * dtype CSEG_F32: the stored value is already normalised (mean / std ignored, chan = {0,1,2});
 * dtype CSEG_U8:  raw bytes; the value is (float(byte) - mean[c]) / std[c] in fp32 (SegDataPreProcessor with
 *                 bgr_to_rgb: chan = {2,1,0} for BGR storage).  HWC uint8 (cv2 / predict_u8): stride_c = 1,
 *                 stride_x = 3, stride_y = 3 W; CHW uint8 (mmengine PackSegInputs): stride_c = img_h * W, stride_x = 1. */
typedef struct cseg_image {
  const void* data;
  int dtype;
  int H, W, img_h;
  long long stride_img, stride_c, stride_y, stride_x;
  int chan[3];
  float mean[3], std[3];
} cseg_image;

/* stand-alone form of the same arithmetic: uint8 HWC BGR image -> float32 [3,H,W] RGB, (x - mean) / std. */
CSEG_API int cseg_preprocess_u8(const uint8_t* img_hwc_bgr, int H, int W, const float mean_rgb_host[3],
                       const float std_rgb_host[3], float* out_chw, void* stream);

/* ---- A2 stem: open_clip/transformer.py:559-576 ------------------------------------------------
 * windows: int32 [n_crops][4] = {y1, x1, h, w} of forward_slide (segmentor.py:418-424); each window
 * is placed at (pad_top, pad_left) inside a zero canvas of crop_h x crop_w (compute_padsize,
 * segmentor.py:427-431,534-546).  out: T [n_crops*gh*gw, ldo], column = c*ps*ps + ky*ps + kx
 * (the flattened conv1.weight order), columns >= 3*ps*ps are written as zero up to ldo. */
CSEG_API int cseg_patchify(const cseg_image* img, const int32_t* windows, int n_crops,
                  int crop_h, int crop_w, int pad_top, int pad_left, int ps, int out_dtype,
                  void* out, int ldo, void* stream);
/* text tower stem (A13/N3; open_clip/model.py:291-293): out[r] = table[idx[r]] + pos[r % L] for r < n_rows
 * (table fp32 [*, width], idx int64 [n_rows], pos fp32 [L, width] or NULL = plain row gather, used for the
 * EOT-token pick of model.py:302-304). */
CSEG_API int cseg_gather_rows(const float* table, const long long* idx, const float* pos, long long n_rows, int L,
                     int width, float* out, void* stream);
/* x[crop*L + t] = (t == 0 ? class_embedding : patch_embed[crop*P + t-1]) + pos[t]   (:565-571) */
CSEG_API int cseg_embed_tokens(const float* patch_embed, const float* class_embedding, const float* pos,
                      int n_crops, int L, int width, float* x, void* stream);
/* the same followed by ln_pre (:574) in one pass over the rows: x = LayerNorm(tokens + pos) with fp32 statistics
 * (width % 4 == 0, 16-byte aligned pointers); equals cseg_embed_tokens + in-place cseg_layernorm bit for bit. */
CSEG_API int cseg_embed_tokens_ln(const float* patch_embed, const float* class_embedding, const float* pos,
                         int n_crops, int L, int width, const float* gamma, const float* beta, float eps,
                         float* x, void* stream);
/* LayerNormFp32 (open_clip/transformer.py:17-23): fp32 statistics, output in out_dtype.
 * `x` and `out` may alias when out_dtype == CSEG_F32. */
CSEG_API int cseg_layernorm(const float* x, int rows, int width, const float* gamma, const float* beta,
                   float eps, int out_dtype, void* out, void* stream);

/* ---- GEMM (K1,K3,K4,K9,K12 of SURVEY 2.2): nn.Linear / 1x1 conv -------------------------------
 * C[M,N] = residual + alpha * act(A[M,K] . B[N,K]^T + bias)      (residual, bias optional)
 * in_dtype CSEG_BF16 -> TMA-fed tcgen05 kernel (fp32 accumulate in TMEM); A,B bf16, K % 8 == 0,
 * lda,ldb % 8 == 0, 16-byte aligned bases.  in_dtype CSEG_F32 -> CUDA-core fp32 verification
 * kernel.  bias fp32 [N]; residual [M, ldr] in res_dtype; C in out_dtype (may alias residual when both
 * have the same dtype). */
CSEG_API int cseg_gemm(int in_dtype, const void* A, int lda, const void* B, int ldb, int M, int N, int K,
              const float* bias, const void* residual, int ldr, int res_dtype, float alpha, int act,
              int out_dtype, void* C, int ldc, void* stream);
/* bf16 A[M,K] . B[N,K]^T restricted to the block diagonal: only output tiles that touch a square diagonal block of
 * block_rows rows / columns are computed (the per-crop Gram matrices of cseg_basis_logits in ONE launch: entries that pair
 * rows of different blocks are left untouched).  Same operand constraints as cseg_gemm(CSEG_BF16). */
CSEG_API int cseg_gemm_blockdiag(const void* A, int lda, const void* B, int ldb, int M, int N, int K, int block_rows,
                        int out_dtype, void* C, int ldc, void* stream);
/* same contract on the CUDA-core kernel for either operand dtype: the on-device cross-check of the
 * tensor-core path used by the tests (never called by the product path). */
CSEG_API int cseg_gemm_reference(int in_dtype, const void* A, int lda, const void* B, int ldb, int M, int N,
                        int K, const float* bias, const void* residual, int ldr, int res_dtype,
                        float alpha, int act, int out_dtype, void* C, int ldc, void* stream);

/* ---- attention (K3, K5, K7) ---------------------------------------------------------------------
 * qkv: T [n_crops*L, 3*width] (= F.linear(x, in_proj_weight, in_proj_bias), :841); heads split as
 * :842-844.  mode selects the weight formula; simmap (fp32 [n_crops, L-1, L-1], may be NULL) is the
 * SimilarityEnhancementModule map added with weight sim_weight and a zero CLS row/column
 * (similarity_enhancement.py:78-124).  out: T [n_crops*L, width] (before out_proj).
 * stats (may be NULL; CSEG_ATTN_STD only): fp32 [n_crops][heads][2][L-1] receiving P[0,1+i] and
 * P[1+i,1+i] per head -- the only entries of the need_weights=True matrix that
 * detect_outliers_by_attention consumes (outlier_suppression.py:46-53).
 * Any L >= 2: bf16, head_dim 64, CSEG_ATTN_STD runs on tcgen05 up to L = 272 (197 / 257 = ViT-B/16 / ViT-L/14 crops),
 * the other modes on mma.sync up to 272, then the CUDA-core kernels (Q / K / V of a head in shared memory up to
 * L = 320, a score row per warp with streamed K / V beyond: whole-image inference, segmentor.py:470-471). */
CSEG_API int cseg_attention(int dtype, const void* qkv, int n_crops, int L, int heads, int head_dim, int mode,
                   const float* simmap, float sim_weight, void* out, float* stats, void* stream);
/* SimilarityEnhancementModule.compute_similarity_map (similarity_enhancement.py:37-66) on the fp32
 * residual stream x [n_crops*L, width] (CLS row skipped): M [n_crops, L-1, L-1] fp32. */
CSEG_API int cseg_simmap(const float* x, int n_crops, int L, int width, float temperature,
                int add_self_similarity, float* simmap, void* stream);
/* the same map (add_self_similarity = true) on the tensor cores: rows are normalised and split into bf16 [hi | lo]
 * (scratch: bf16 [n_crops*L, 2*width], 16-byte aligned), then ONE block-diagonal tcgen05 GEMM accumulates
 * hi.hi + hi.lo + lo.hi in fp32 and stores the per-crop blocks compactly.  |difference to cseg_simmap| <= 2e-5 / temperature.
 * width % 64 == 0.
 * layout 0: M [n_crops, L-1, L-1] as cseg_simmap.
 * layout 1 (L <= CSEG_SIMT_COLS_MAX): the zero-padded map of enhance_attention (similarity_enhancement.py:104-107) in the
 *   order the tcgen05 attention reads it: fp32 [n_crops][ceil(C / 32)][C][32] indexed [crop][i / 32][j][i % 32] for token
 *   indices i, j (CLS = 0), with C = CSEG_SIMT_COLS_FOR(L) key columns (208 for ViT-B/16 crops, 272 for ViT-L/14 crops).
 *   Only i, j >= 1 are written: the caller zero-fills the buffer once (CSEG_SIMT_FLOATS_FOR(L) per crop). */
#define CSEG_SIMT_COLS 208
#define CSEG_SIMT_COLS_MAX 272
#define CSEG_SIMT_COLS_FOR(L) ((L) <= CSEG_SIMT_COLS ? CSEG_SIMT_COLS : CSEG_SIMT_COLS_MAX)
#define CSEG_SIMT_FLOATS_FOR(L) ((CSEG_SIMT_COLS_FOR(L) + 31) / 32 * CSEG_SIMT_COLS_FOR(L) * 32)
#define CSEG_SIMT_FLOATS CSEG_SIMT_FLOATS_FOR(1)
CSEG_API int cseg_simmap_tc(const float* x, int n_crops, int L, int width, float temperature, void* scratch,
                   float* simmap, int layout, void* stream);
/* model_type 'Experimental' final-block attention (open_clip/transformer.py:897-903) on tcgen05 with the similarity
 * map in layout 1 of cseg_simmap_tc (simmap_t may be NULL: no enhancement): out = softmax(softmax((k k^T + q q^T) / 8)
 * + sim_weight * M_pad) v.  bf16 qkv / out, head_dim 64, 17 <= L <= CSEG_SIMT_COLS_MAX. */
CSEG_API int cseg_attention_experimental_tc(const void* qkv, int n_crops, int L, int heads, const float* simmap_t,
                                   float sim_weight, void* out, void* stream);
/* detect_outliers_by_attention + OutlierSuppressionModule.mean_interpolation
 * (outlier_suppression.py:15-61,115-214): y [n_crops*L, width] fp32 -> y_out (same shape, OUT OF PLACE,
 * CLS row copied).  grid x grid patches, stats from cseg_attention.  plan: int32 workspace of
 * n_crops * (25*top_k + L-1) words.  outlier_idx (int32 [n_crops, top_k], may be NULL) receives the
 * top-k order. */
CSEG_API int cseg_outlier_suppress(const float* y, float* y_out, int n_crops, int L, int width, int grid,
                          const float* stats, int heads, int top_k, float contamination_temp,
                          int32_t* plan, int32_t* outlier_idx, void* stream);
/* forward_feature head (segmentor.py:309-336): tok fp32 [n_crops*L, D] = ln_post(x) @ proj.
 * cls_unit[crop] = tok[crop*L] / |.|;  feats[crop, p] = f - factor * cos(f, cls) * cls_unit
 * written as T [n_crops*rows_per_crop, ldf] (channel-last patch grid); rows_per_crop = 0 means L-1, a larger
 * value appends zero rows to every crop (16-byte aligned per-crop blocks for cseg_basis_logits). */
CSEG_API int cseg_cls_debias(const float* tok, int n_crops, int L, int D, float factor, int out_dtype,
                    void* feats, int ldf, int rows_per_crop, float* cls_unit, void* stream);

/* ---- JBU upsampler (K11): simfeatup_dev/upsamplers.py:202-325 ---------------------------------
 * guidance for one stage: adaptive_avg_pool2d of each crop to (gh, gw) (:316) -> fp32 [n,gh,gw,4]
 * (RGB + 0 pad). */
CSEG_API int cseg_jbu_guidance(const cseg_image* img, const int32_t* windows, int n_crops,
                      int crop_h, int crop_w, int pad_top, int pad_left, int gh, int gw,
                      float* guid, void* stream);
/* the two calls above/below in one kernel for the bf16 pipeline: guid (fp32 [n,gh,gw,4]) and proj (CSEG_F16
 * [n,gh,gw,32]) of one stage from the image; the hidden GELU uses the tanh form (|err| <= 4.8e-4, the size of the
 * fp16 rounding the reference's autocast applies to these activations). */
CSEG_API int cseg_jbu_guidance_proj(const cseg_image* img, const int32_t* windows, int n_crops,
                           int crop_h, int crop_w, int pad_top, int pad_left, int gh, int gw, int key_dim,
                           const float* w0, const float* b0, const float* w3, const float* b3, float* guid,
                           int proj_dtype, void* proj, void* stream);
/* range_proj (:209-214): conv1x1(3->kd) . GELU . conv1x1(kd->kd); proj [n,gh,gw,kd] in proj_dtype:
 * CSEG_F32 (verification mode) or CSEG_F16 (tensor-core range kernel; the reference computes these in
 * fp16 under autocast, segmentor.py:370). */
CSEG_API int cseg_jbu_range_proj(const float* guid, int n_pix, int key_dim, const float* w0, const float* b0,
                        const float* w3, const float* b3, int proj_dtype, void* proj, void* stream);
/* get_range_kernel x get_spatial_kernel, renormalised (:230-251,258-262).  kern: T rows of stride ldk;
 * columns [0,d*d) = combined kernel, [d*d, d*d+3) = guidance RGB (the fixup_proj input order, :264),
 * [d*d+3, kwidth) = zero; columns >= kwidth are not touched (the engine keeps the fix-up hidden layer in
 * the other half of the same rows). */
CSEG_API int cseg_jbu_range_kernel(int proj_dtype, const void* proj, const float* guid, int n_crops, int gh,
                          int gw, int key_dim, int radius, float range_temp, float sigma_spatial,
                          int out_dtype, void* kern, int kwidth, int ldk, void* stream);
/* bicubic x2 (align_corners=False, a=-0.75) + reflect pad + adaptive conv (:268-274, semantics of
 * adaptive_conv_py_simple :14-25).  src T [n, h, w, C] channel-last -> dst T [n, 2h, 2w, C];
 * kern T [n*2h*2w, ldk] (first d*d columns used); hr_scratch: T [n*2h*2w*C] workspace. */
CSEG_API int cseg_jbu_apply(int dtype, const void* src, int n_crops, int h, int w, int C, const void* kern,
                   int ldk, int radius, void* dst, void* hr_scratch, void* stream);

/* ---- JBU kernel generation shared across overlapping crops (bf16 path) ---------------------------------------
 * guidance pooling, range projection, range kernel, kernel fix-up and the bicubic-folded composite kernels depend on
 * the IMAGE, not the crop, for every pixel further than radius + 6 from the crop border (reflect padding and bicubic
 * border clamps, upsamplers.py:233,268-269).  With overlapping windows they are therefore computed once per image
 * pixel ("image level": one region = the whole canvas, through the plain entry points above with n_crops = 1) plus,
 * per crop, for a border frame: the top / bottom fb rows and the left / right 16 columns, stored compactly
 * (cseg_jbu_share_rows(gh, gw, fb) rows per crop: top strip, bottom strip, then 32 pixels per remaining row).
 * Requires full-size windows whose origins are multiples of 16 image pixels.  fb = CSEG_JBU_FB_RANGE for the
 * range-kernel / fix-up tensors, CSEG_JBU_FB_COMP for the composite kernels (bicubic clamps reach further). */
#define CSEG_JBU_FB_RANGE 8
#define CSEG_JBU_FB_COMP 12
typedef struct cseg_jbu_share {
  const int32_t* windows; /* device int32 [n_crops][4] = {y1, x1, h, w} (canvas pixels), as for cseg_patchify */
  int shift;              /* log2(image pixels per pixel of this stage): crop origin at this stage = (y1 >> shift, x1 >> shift) */
  int pitch;              /* pixels per row of the image-level buffers of this stage (canvas W >> shift) */
} cseg_jbu_share;
CSEG_API int cseg_jbu_share_rows(int gh, int gw, int fb);
/* range kernel (as cseg_jbu_range_kernel, fp16 projections -> bf16) of the border frames only: proj_img CSEG_F16
 * [ih*iw, 32] and guid_img fp32 [ih*iw, 4] are IMAGE-LEVEL; kern_border bf16 [n_crops * rows(gh, gw, FB_RANGE), ldk]. */
CSEG_API int cseg_jbu_range_kernel_border(const void* proj_img, const float* guid_img, const cseg_jbu_share* share,
                                 int n_crops, int gh, int gw, int radius, float range_temp, float sigma_spatial,
                                 void* kern_border, int kwidth, int ldk, void* stream);
/* composite (bicubic-folded) kernels of every image-level pixel, valid for crop-interior pixels: kern_img bf16
 * [ih*iw, ldk] (after the fix-up) -> kc_img bf16 [ih*iw, 128].  gh x gw = the crop region at this stage (selects the
 * interior bicubic phase tables); tabs_scratch: (gh + gw) * 512 bytes. */
CSEG_API int cseg_jbu_composite_image(const void* kern_img, int ldk, int ih, int iw, int gh, int gw, int radius,
                             void* kc_img, void* tabs_scratch, void* stream);
/* cseg_jbu_apply for n crops with shared kernels: interior pixels read kc_img at the crop's origin, border-frame
 * pixels get their composite kernels here from kern_border (frame FB_RANGE) / kern_img.  src bf16 [n, h, w, C] ->
 * dst bf16 [n, 2h, 2w, C].  scratch: n * rows(2h, 2w, FB_COMP) * 256 + (2h + 2w) * 512 bytes. */
CSEG_API int cseg_jbu_apply_shared(const void* src, int n_crops, int h, int w, int C, const void* kern_border,
                          const void* kern_img, const void* kc_img, const cseg_jbu_share* share, int ldk, int radius,
                          void* dst, void* scratch, void* stream);

/* ---- A10: L2-normalise + cosine logits, segmentor.py:374-375,378-379 --------------------------
 * feats T [rows, ldf] (D used) ; text fp32 [Q, D] ; logits fp32 [n_crops, Q, hw] with rows =
 * n_crops*hw.  cls_logit_bias (fp32 [n_crops, Q], may be NULL) is added (cls_token_lambda term). */
CSEG_API int cseg_norm_sim(int dtype, const void* feats, int ldf, int n_crops, int hw, int D,
                  const float* text, int Q, const float* cls_logit_bias, float* logits, void* stream);

/* K12 + K13 fused: final 1x1 conv of the upsampler (upsamplers.py:325) + normalise + cosine logits:
 *   out = y + alpha * (y . W^T + bias);  logits[crop, q, pix] = <out/|out|, text[q]> (+ cls bias).
 * y T [n_crops*hw, ldy] channel-last, W T [C, ldw].  In bf16 (C % 128 == 0, C <= 512, Q <= 16) this is ONE
 * tcgen05 kernel whose epilogue keeps `out` on chip; otherwise cseg_gemm + cseg_norm_sim through `scratch`
 * (T [n_crops*hw*C], may be NULL when the fused kernel applies). */
CSEG_API int cseg_fixup_norm_sim(int dtype, const void* y, int ldy, const void* W, int ldw, int n_crops, int hw,
                        int C, const float* bias, float alpha, const float* text, int Q,
                        const float* cls_logit_bias, float* logits, void* scratch, void* stream);

/* JBU kernel fix-up in one pass (upsamplers.py:218-223,258-262), bf16:
 *   out[p, :] = k[p, :] + (W3s . gelu(W0 . k[p, :] + b0) + b3s)      W3s = 0.1 * fixup_proj.3.weight, b3s = 0.1 * bias
 * k, out bf16 [M, ldk] (row strides lda, ldo); W0, W3s bf16 [ldk, ldk] row-major (row strides ldw0, ldw3), zero padded
 * beyond the (2r+1)^2 (+3 guidance) used entries; ldk = 64 or 128.  Two chained tcgen05 GEMMs per 128-row panel; the
 * hidden activations stay in shared memory.  Equivalent to two cseg_gemm calls (GELU epilogue, then residual). */
CSEG_API int cseg_jbu_kernel_fixup(int dtype, const void* k, int lda, const void* W0, int ldw0, const float* b0,
                          const void* W3s, int ldw3, const float* b3s, int M, int ldk, void* out, int ldo,
                          void* stream);

/* K11..K13 in basis form (bf16 fast path; same result as cseg_jbu_apply x4 + cseg_fixup_norm_sim).
 * The JBU stack (upsamplers.py:269-274,320-325) is linear in its source and treats every channel alike, so
 * upsampling the identity (one channel per low-resolution token: src[crop, k, :] = e_k) gives coefficients
 * s[p, k] with  out[p] = sum_k s[p, k] g[crop, k] + b,  g = tokens after the final 1x1 conv, and
 *   |out|^2 = s^T (g g^T) s + 2 s.(g b) + b.b        <out, text[q]> = s.(g text[q]) + b.text[q].
 * s    bf16 [n_crops*hw, lds], Cb valid columns (>= 16*ceil(T/16); columns >= T are zero), hw % 128 == 0
 * gram bf16 [n_crops*tstride, ldg]: gram[crop*tstride + j, crop*tstride + k] = <g[crop, j], g[crop, k]> (only the
 *      diagonal blocks are read; entries that pair a token with padding or with another crop may hold anything finite)
 * aux  bf16 [16 or 32, ldg] (32 rows when Q + 1 > 16): aux[q, crop*tstride + k] = <g[crop, k], text[q]> (q < Q),
 *      row Q = <g[crop, k], b>
 * consts fp32 [Q+1]: <b, text[q]>, then <b, b>.    16*ceil(T/16) + (Q < 16 ? 16 : 32) <= 256, Q <= 31, tstride % 8 == 0
 * (TMA box origins must be 16-byte aligned).
 * logits fp32 [n_crops, Q, hw] = cosine (+ cls_logit_bias[crop, q], segmentor.py:378-379). */
CSEG_API int cseg_basis_logits(int dtype, const void* s, int lds, int Cb, int n_crops, int hw, int T, int tstride,
                      const void* gram, const void* aux, int ldg, const float* consts, int Q,
                      const float* cls_logit_bias, float* logits, void* stream);

/* ---- A11 + A12: forward_slide accumulation + postprocess_result, segmentor.py:413-449,475-499 --
 * crop_logits fp32 [n_crops, Q, lh, lw]; when (lh,lw) != (crop_h,crop_w) each crop is first
 * resized bilinearly (align_corners=False) to crop_h x crop_w (segmentor.py:388-391).  The window
 * of crop i covers rows [y1, y1+h), cols [x1, x1+w) of the H x W canvas and reads the crop at offset
 * (pad_top, pad_left).  Sums over covering windows, divides by the count, resizes to out_h x out_w
 * (:448-449), then x logit_scale, softmax over Q, per-class max over synonym queries
 * (query_idx int32 [Q]), argmax (lowest index wins ties), prob < prob_thd -> bg_idx.
 * labels uint8 [out_h,out_w]; probs (fp32 [K,out_h,out_w]) and avg_logits (fp32 [Q,H,W]) optional. */
CSEG_API int cseg_accum_argmax(const float* crop_logits, int n_crops, int Q, int lh, int lw, int crop_h,
                      int crop_w, int pad_top, int pad_left, const int32_t* windows, int H, int W,
                      int out_h, int out_w, const int32_t* query_idx, int K, float logit_scale,
                      float prob_thd, int bg_idx, uint8_t* labels, float* probs, float* avg_logits,
                      void* stream);

/* ---- N4 output side: segmentor.py:513-531,568-608 ---------------------------------------------------------
 * colourised mask: out_bgr[i] = lut[min(labels[i], n_lut-1)] (lut uint8 [n_lut][3], rows already in the channel order
 * to be written); confidence heat-map: out_bgr[i] = lut256[uint8(clip(nan_to_num(max_k probs[k][i]), 0, 1) * 255)]
 * (lut256 = the 256-entry colour map, e.g. cv2.COLORMAP_JET).  Both write uint8 [n][3] images ready for cv2.imwrite. */
CSEG_API int cseg_colorize(const uint8_t* labels, long long n, const uint8_t* lut, int n_lut, uint8_t* out_bgr, void* stream);
CSEG_API int cseg_heatmap(const float* probs, int K, long long n, const uint8_t* lut256, uint8_t* out_bgr, void* stream);

/* ---- K17: mmseg IoUMetric.intersect_and_union (mmsegmentation 1.2.2, external) -----------------
 * hist int64 [3][K] += {intersect, pred, label} pixel counts; label == ignore_index skipped. */
CSEG_API int cseg_iou_hist(const uint8_t* pred, const uint8_t* label, long long n, int K, int ignore_index,
                  long long* hist, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CLIPSEG_H_ */
