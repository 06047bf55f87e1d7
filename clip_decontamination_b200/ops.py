"""Thin torch-tensor wrappers over the C ABI (pointers + sizes; no torch types cross the boundary).

PyTorch is used here for device memory and streams only.  Every wrapper validates device /
dtype / contiguity, then forwards ``tensor.data_ptr()`` and the current CUDA stream.
"""
import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import lib, check, F32, BF16, F16, U8

_DT = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}

MEAN_RGB = (122.771, 116.746, 104.094)        # segmentor.py:64-67 (SegDataPreProcessor, after bgr_to_rgb)
STD_RGB = (68.501, 66.632, 70.323)


def _stream():
    # the current stream of the CURRENT device: the engines wrap their entry points in torch.cuda.device(engine.device),
    # so a model built with device='cuda:1' launches on cuda:1 whatever the caller's current device is
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Image:
    """``cseg_image`` over a torch tensor: how the kernels that read the input (patchify, JBU guidance) address it.
    A batch of B equally sized images is ONE canvas of B*h rows (image b = rows [b*h, (b+1)*h)); no copy, no
    separate preprocessing pass -- uint8 inputs are normalised on load ((x - mean) / std, segmentor.py:64-67).

      Image.normalised(t)   t fp32 [3,H,W] or [B,3,H,W]: already normalised RGB (what predict() receives)
      Image.u8(t, 'hwc')    t uint8 [H,W,3] / [B,H,W,3], BGR (cv2 / predict_u8)
      Image.u8(t, 'chw')    t uint8 [3,H,W] / [B,3,H,W], BGR (mmengine PackSegInputs -> test_step)
    """

    def __init__(self, t: torch.Tensor, kind: str, B: int, h: int, W: int, strides, chan=(0, 1, 2),
                 mean=MEAN_RGB, std=STD_RGB):
        if not t.is_cuda:
            raise _lib.ClipSegError('libclipseg needs CUDA tensors (there is no CPU path)')
        if not t.is_contiguous():
            raise _lib.ClipSegError('libclipseg needs contiguous tensors')
        self.tensor, self.kind, self.B, self.h, self.W, self.H = t, kind, B, h, W, B * h
        d = _lib.CsegImage()
        d.data = t.data_ptr()
        d.dtype = U8 if t.dtype == torch.uint8 else F32
        d.H, d.W, d.img_h = B * h, W, h
        d.stride_img, d.stride_c, d.stride_y, d.stride_x = strides
        d.chan = (C.c_int * 3)(*chan)
        d.mean = (C.c_float * 3)(*mean)
        d.std = (C.c_float * 3)(*std)
        self.desc = d

    @property
    def device(self):
        return self.tensor.device

    @classmethod
    def normalised(cls, t: torch.Tensor):
        assert t.dtype == torch.float32 and t.dim() in (3, 4)
        if t.dim() == 3:
            _, H, W = t.shape
            return cls(t, 'f32', 1, H, W, (0, H * W, W, 1))
        B, _, H, W = t.shape
        return cls(t, 'f32', B, H, W, (3 * H * W, H * W, W, 1))

    @classmethod
    def u8(cls, t: torch.Tensor, layout: str = 'hwc', mean=MEAN_RGB, std=STD_RGB, bgr_to_rgb: bool = True):
        assert t.dtype == torch.uint8 and t.dim() in (3, 4) and layout in ('hwc', 'chw')
        if t.dim() == 3:
            t = t.unsqueeze(0)
        chan = (2, 1, 0) if bgr_to_rgb else (0, 1, 2)
        if layout == 'hwc':
            B, H, W, _ = t.shape
            return cls(t, 'u8hwc', B, H, W, (3 * H * W, 1, 3 * W, 3), chan, mean, std)
        B, _, H, W = t.shape
        return cls(t, 'u8chw', B, H, W, (3 * H * W, H * W, W, 1), chan, mean, std)


def _ptr(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.ClipSegError('libclipseg needs CUDA tensors (there is no CPU path)')
    if not t.is_contiguous():
        raise _lib.ClipSegError('libclipseg needs contiguous tensors')
    return C.c_void_p(t.data_ptr())


def _dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise _lib.ClipSegError(f'unsupported dtype {t.dtype} (float32 / bfloat16)')


def preprocess_u8(img_hwc_bgr: torch.Tensor, mean, std, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    H, W, _ = img_hwc_bgr.shape
    assert img_hwc_bgr.dtype == torch.uint8
    if out is None:
        out = torch.empty((3, H, W), dtype=torch.float32, device=img_hwc_bgr.device)
    check(lib.cseg_preprocess_u8(_ptr(img_hwc_bgr), H, W, (C.c_float * 3)(*mean), (C.c_float * 3)(*std),
                                 _ptr(out), _stream()))
    return out


def _image(img) -> Image:
    return img if isinstance(img, Image) else Image.normalised(img)


def patchify(img, windows: torch.Tensor, crop_h: int, crop_w: int, pad_top: int, pad_left: int,
             ps: int, out: torch.Tensor):
    assert windows.dtype == torch.int32
    check(lib.cseg_patchify(C.byref(_image(img).desc), _ptr(windows), windows.shape[0], crop_h, crop_w, pad_top,
                            pad_left, ps, _dt(out), _ptr(out), out.shape[1], _stream()))
    return out


def gather_rows(table: torch.Tensor, idx: torch.Tensor, pos: Optional[torch.Tensor], L: int, out: torch.Tensor):
    """out[r] = table[idx[r]] (+ pos[r % L]): token embedding + positional embedding of the text tower, and the EOT pick."""
    assert table.dtype == torch.float32 and idx.dtype == torch.int64 and out.dtype == torch.float32
    check(lib.cseg_gather_rows(_ptr(table), _ptr(idx), _ptr(pos), idx.numel(), L, table.shape[1], _ptr(out), _stream()))
    return out


def embed_tokens(pe, cls_emb, pos, n_crops, L, width, x):
    check(lib.cseg_embed_tokens(_ptr(pe), _ptr(cls_emb), _ptr(pos), n_crops, L, width, _ptr(x), _stream()))
    return x


def embed_tokens_ln(pe, cls_emb, pos, n_crops, L, width, gamma, beta, x, eps: float = 1e-5):
    """embed_tokens followed by ln_pre in one pass (same values)."""
    check(lib.cseg_embed_tokens_ln(_ptr(pe), _ptr(cls_emb), _ptr(pos), n_crops, L, width, _ptr(gamma), _ptr(beta), eps,
                                   _ptr(x), _stream()))
    return x


def layernorm(x: torch.Tensor, gamma, beta, out: torch.Tensor, eps: float = 1e-5):
    assert x.dtype == torch.float32
    rows, width = x.shape
    check(lib.cseg_layernorm(_ptr(x), rows, width, _ptr(gamma), _ptr(beta), eps, _dt(out), _ptr(out), _stream()))
    return out


def gemm(A: torch.Tensor, B: torch.Tensor, out: torch.Tensor, bias=None, residual=None, alpha: float = 1.0,
         act: int = 0, M: Optional[int] = None, N: Optional[int] = None, K: Optional[int] = None,
         reference: bool = False):
    """out[M,N] = residual + alpha * act(A[M,K] @ B[N,K]^T + bias).  2-D row-major views; strides taken
    from the tensors, so column-padded buffers work."""
    assert A.dim() == 2 and B.dim() == 2 and out.dim() == 2 and A.dtype == B.dtype
    assert A.stride(1) == 1 and B.stride(1) == 1 and out.stride(1) == 1
    M = A.shape[0] if M is None else M
    N = B.shape[0] if N is None else N
    K = A.shape[1] if K is None else K
    fn = lib.cseg_gemm_reference if reference else lib.cseg_gemm
    ldr = residual.stride(0) if residual is not None else 0
    rdt = _dt(residual) if residual is not None else F32
    if residual is not None:
        assert residual.stride(1) == 1
    if bias is not None:
        assert bias.dtype == torch.float32
    check(fn(_dt(A), C.c_void_p(A.data_ptr()), A.stride(0), C.c_void_p(B.data_ptr()), B.stride(0), M, N, K,
             _ptr(bias), C.c_void_p(residual.data_ptr()) if residual is not None else None, ldr, rdt, alpha, act,
             _dt(out), C.c_void_p(out.data_ptr()), out.stride(0), _stream()))
    return out


def gemm_blockdiag(A: torch.Tensor, B: torch.Tensor, out: torch.Tensor, block_rows: int):
    """out = A @ B^T on the block diagonal only (square blocks of block_rows); other entries of `out` are not written."""
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16 and A.stride(1) == 1 and B.stride(1) == 1 and out.stride(1) == 1
    check(lib.cseg_gemm_blockdiag(C.c_void_p(A.data_ptr()), A.stride(0), C.c_void_p(B.data_ptr()), B.stride(0), A.shape[0],
                                  B.shape[0], A.shape[1], block_rows, _dt(out), C.c_void_p(out.data_ptr()), out.stride(0),
                                  _stream()))
    return out


def attention(qkv: torch.Tensor, n_crops: int, L: int, heads: int, head_dim: int, mode: int, out: torch.Tensor,
              simmap=None, sim_weight: float = 1.0, stats=None):
    check(lib.cseg_attention(_dt(qkv), _ptr(qkv), n_crops, L, heads, head_dim, mode, _ptr(simmap), sim_weight,
                             _ptr(out), _ptr(stats), _stream()))
    return out


SIMT_COLS, SIMT_COLS_MAX = 208, 272               # CSEG_SIMT_COLS / CSEG_SIMT_COLS_MAX of include/clipseg.h


def simt_cols(L: int) -> int:
    """key columns of the transposed similarity map for crops of L tokens (CSEG_SIMT_COLS_FOR)"""
    return SIMT_COLS if L <= SIMT_COLS else SIMT_COLS_MAX


def simt_floats(L: int) -> int:
    """floats per crop of the transposed similarity map (CSEG_SIMT_FLOATS_FOR)"""
    c = simt_cols(L)
    return (c + 31) // 32 * c * 32


SIMT_FLOATS = simt_floats(1)


def simmap(x: torch.Tensor, n_crops: int, L: int, width: int, out: torch.Tensor, temperature: float = 1.0,
           add_self_similarity: bool = True, scratch: Optional[torch.Tensor] = None, transposed: bool = False):
    """scratch (bf16 [n_crops*L, 2*width]) selects the tensor-core form (hi/lo bf16 split, fp32-grade products).
    transposed: `out` is the zero-initialised padded map [n_crops, simt_floats(L)] in the order attention_experimental_tc
    reads it (layout 1 of cseg_simmap_tc); otherwise [n_crops, L-1, L-1]."""
    assert x.dtype == torch.float32 and out.dtype == torch.float32
    if scratch is not None:
        assert add_self_similarity and scratch.dtype == torch.bfloat16 and scratch.numel() >= n_crops * L * 2 * width
        assert out.numel() >= (n_crops * simt_floats(L) if transposed else n_crops * (L - 1) * (L - 1))
        check(lib.cseg_simmap_tc(_ptr(x), n_crops, L, width, temperature, _ptr(scratch), _ptr(out), int(transposed), _stream()))
        return out
    assert not transposed
    check(lib.cseg_simmap(_ptr(x), n_crops, L, width, temperature, int(add_self_similarity), _ptr(out), _stream()))
    return out


def attention_experimental_tc(qkv: torch.Tensor, n_crops: int, L: int, heads: int, out: torch.Tensor, simmap_t=None,
                              sim_weight: float = 1.0):
    """Final-block 'Experimental' attention on tcgen05 (head_dim 64, L <= SIMT_COLS_MAX); simmap_t: simmap(..., transposed=True)."""
    assert qkv.dtype == torch.bfloat16 and out.dtype == torch.bfloat16 and qkv.shape[1] == 3 * heads * 64
    check(lib.cseg_attention_experimental_tc(_ptr(qkv), n_crops, L, heads, _ptr(simmap_t), sim_weight, _ptr(out), _stream()))
    return out


def outlier_suppress(y, y_out, n_crops, L, width, grid, stats, heads, top_k, contamination_temp, plan, outlier_idx=None):
    assert plan.dtype == torch.int32 and plan.numel() >= n_crops * (25 * top_k + L - 1)
    check(lib.cseg_outlier_suppress(_ptr(y), _ptr(y_out), n_crops, L, width, grid, _ptr(stats), heads, top_k,
                                    contamination_temp, _ptr(plan), _ptr(outlier_idx), _stream()))
    return y_out


def cls_debias(tok, n_crops, L, D, factor, feats, cls_unit=None, rows_per_crop=0):
    check(lib.cseg_cls_debias(_ptr(tok), n_crops, L, D, factor, _dt(feats), _ptr(feats), feats.shape[-1],
                              rows_per_crop, _ptr(cls_unit), _stream()))
    return feats


def jbu_guidance(img, windows, crop_h, crop_w, pad_top, pad_left, gh, gw, out):
    check(lib.cseg_jbu_guidance(C.byref(_image(img).desc), _ptr(windows), windows.shape[0], crop_h, crop_w, pad_top,
                                pad_left, gh, gw, _ptr(out), _stream()))
    return out


def jbu_guidance_proj(img, windows, crop_h, crop_w, pad_top, pad_left, gh, gw, w0, b0, w3, b3, guid, proj):
    """guidance + range projection of one stage in one kernel (bf16 pipeline; proj fp16)."""
    check(lib.cseg_jbu_guidance_proj(C.byref(_image(img).desc), _ptr(windows), windows.shape[0], crop_h, crop_w, pad_top,
                                     pad_left, gh, gw, 32, _ptr(w0), _ptr(b0), _ptr(w3), _ptr(b3), _ptr(guid), _dt(proj),
                                     _ptr(proj), _stream()))
    return guid, proj


def jbu_range_proj(guid, n_pix, w0, b0, w3, b3, proj):
    check(lib.cseg_jbu_range_proj(_ptr(guid), n_pix, 32, _ptr(w0), _ptr(b0), _ptr(w3), _ptr(b3), _dt(proj), _ptr(proj),
                                  _stream()))
    return proj


def jbu_range_kernel(proj, guid, n_crops, gh, gw, radius, range_temp, sigma_spatial, kern):
    """kern: 2-D (possibly column-sliced) view [n*gh*gw, kwidth]; its row stride is passed as ldk."""
    assert kern.dim() == 2 and kern.stride(1) == 1
    check(lib.cseg_jbu_range_kernel(_dt(proj), _ptr(proj), _ptr(guid), n_crops, gh, gw, 32, radius, range_temp, sigma_spatial,
                                    _dt(kern), C.c_void_p(kern.data_ptr()), kern.shape[1], kern.stride(0), _stream()))
    return kern


def jbu_apply(src, n_crops, h, w, Cc, kern, radius, dst, hr_scratch):
    check(lib.cseg_jbu_apply(_dt(src), _ptr(src), n_crops, h, w, Cc, _ptr(kern), kern.shape[-1], radius, _ptr(dst),
                             _ptr(hr_scratch), _stream()))
    return dst


# ---- JBU kernels shared across overlapping crops (include/clipseg.h "shared kernel generation") -------------------------
FB_RANGE, FB_COMP = 8, 12                  # CSEG_JBU_FB_RANGE / CSEG_JBU_FB_COMP


def jbu_share_rows(gh: int, gw: int, fb: int) -> int:
    return int(lib.cseg_jbu_share_rows(gh, gw, fb))


def _share(windows: torch.Tensor, shift: int, pitch: int):
    assert windows.dtype == torch.int32 and windows.is_contiguous()
    sh = _lib.CsegJbuShare()
    sh.windows, sh.shift, sh.pitch = windows.data_ptr(), shift, pitch
    return sh


def jbu_range_kernel_border(proj_img, guid_img, windows, shift, pitch, n_crops, gh, gw, radius, range_temp,
                            sigma_spatial, kern_border):
    """Range kernel of the border frames of n crops from the image-level projections (cseg_jbu_range_kernel_border)."""
    assert kern_border.dim() == 2 and kern_border.stride(1) == 1
    assert kern_border.shape[0] >= n_crops * jbu_share_rows(gh, gw, FB_RANGE)
    check(lib.cseg_jbu_range_kernel_border(_ptr(proj_img), _ptr(guid_img), C.byref(_share(windows, shift, pitch)), n_crops,
                                           gh, gw, radius, range_temp, sigma_spatial, C.c_void_p(kern_border.data_ptr()),
                                           kern_border.shape[1], kern_border.stride(0), _stream()))
    return kern_border


def jbu_composite_image(kern_img, ih, iw, gh, gw, radius, kc_img, tabs):
    check(lib.cseg_jbu_composite_image(_ptr(kern_img), kern_img.shape[-1], ih, iw, gh, gw, radius, _ptr(kc_img), _ptr(tabs),
                                       _stream()))
    return kc_img


def jbu_apply_shared(src, n_crops, h, w, Cc, kern_border, kern_img, kc_img, windows, shift, pitch, radius, dst, scratch):
    check(lib.cseg_jbu_apply_shared(_ptr(src), n_crops, h, w, Cc, _ptr(kern_border), _ptr(kern_img), _ptr(kc_img),
                                    C.byref(_share(windows, shift, pitch)), kern_img.shape[-1], radius, _ptr(dst),
                                    _ptr(scratch), _stream()))
    return dst


def norm_sim(feats, ldf, n_crops, hw, D, text, logits, cls_logit_bias=None):
    check(lib.cseg_norm_sim(_dt(feats), _ptr(feats), ldf, n_crops, hw, D, _ptr(text), text.shape[0],
                            _ptr(cls_logit_bias), _ptr(logits), _stream()))
    return logits


def fixup_norm_sim(y, W, n_crops, hw, Cc, bias, alpha, text, logits, cls_logit_bias=None, scratch=None):
    check(lib.cseg_fixup_norm_sim(_dt(y), _ptr(y), y.stride(0), _ptr(W), W.stride(0), n_crops, hw, Cc, _ptr(bias), alpha,
                                  _ptr(text), text.shape[0], _ptr(cls_logit_bias), _ptr(logits), _ptr(scratch), _stream()))
    return logits


def jbu_kernel_fixup(k, W0, b0, W3s, b3s, out):
    """out = k + W3s . gelu(W0 . k + b0) + b3s per row (include/clipseg.h: cseg_jbu_kernel_fixup); 2-D views."""
    M, ldk = k.shape
    assert out.shape == k.shape and W0.shape[0] == ldk and W3s.shape[0] == ldk
    for t in (k, W0, W3s, out):                      # row-strided 2-D views are fine (strides are passed)
        assert t.is_cuda and t.dim() == 2 and t.stride(1) == 1 and t.dtype == k.dtype
    vp = lambda t: C.c_void_p(t.data_ptr())
    check(lib.cseg_jbu_kernel_fixup(_dt(k), vp(k), k.stride(0), vp(W0), W0.stride(0), _ptr(b0), vp(W3s),
                                    W3s.stride(0), _ptr(b3s), M, ldk, vp(out), out.stride(0), _stream()))
    return out


def basis_logits(s, Cb, n_crops, hw, T, tstride, gram, aux, consts, Q, logits, cls_logit_bias=None):
    """Cosine logits from JBU basis coefficients (include/clipseg.h: cseg_basis_logits)."""
    assert gram.stride(0) == aux.stride(0)
    check(lib.cseg_basis_logits(_dt(s), _ptr(s), s.stride(0), Cb, n_crops, hw, T, tstride, _ptr(gram), _ptr(aux),
                                gram.stride(0), _ptr(consts), Q, _ptr(cls_logit_bias), _ptr(logits), _stream()))
    return logits


def accum_argmax(crop_logits, windows, crop_h, crop_w, pad_top, pad_left, H, W, out_h, out_w, query_idx, K,
                 logit_scale, prob_thd, bg_idx, labels, probs=None, avg_logits=None):
    n, Q, lh, lw = crop_logits.shape
    assert crop_logits.dtype == torch.float32 and labels.dtype == torch.uint8
    check(lib.cseg_accum_argmax(_ptr(crop_logits), n, Q, lh, lw, crop_h, crop_w, pad_top, pad_left, _ptr(windows),
                                H, W, out_h, out_w, _ptr(query_idx), K, logit_scale, prob_thd, bg_idx,
                                _ptr(labels), _ptr(probs), _ptr(avg_logits), _stream()))
    return labels


def iou_hist(pred, label, K, hist, ignore_index=255):
    assert pred.dtype == torch.uint8 and label.dtype == torch.uint8 and hist.dtype == torch.int64
    check(lib.cseg_iou_hist(_ptr(pred), _ptr(label), pred.numel(), K, ignore_index, _ptr(hist), _stream()))
    return hist


def colorize(labels: torch.Tensor, palette: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """uint8 labels [H,W] + uint8 palette [n,3] -> uint8 image [H,W,3] (palette rows in the order to be written)."""
    assert labels.dtype == torch.uint8 and palette.dtype == torch.uint8 and palette.shape[1] == 3
    if out is None:
        out = torch.empty(tuple(labels.shape) + (3,), dtype=torch.uint8, device=labels.device)
    check(lib.cseg_colorize(_ptr(labels), labels.numel(), _ptr(palette), palette.shape[0], _ptr(out), _stream()))
    return out


def heatmap(probs: torch.Tensor, lut256: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 probabilities [K,H,W] + uint8 colour map [256,3] -> uint8 image [H,W,3] of the per-pixel max probability."""
    assert probs.dtype == torch.float32 and lut256.dtype == torch.uint8 and tuple(lut256.shape) == (256, 3)
    K, H, W = probs.shape
    if out is None:
        out = torch.empty((H, W, 3), dtype=torch.uint8, device=probs.device)
    check(lib.cseg_heatmap(_ptr(probs), K, H * W, _ptr(lut256), _ptr(out), _stream()))
    return out
