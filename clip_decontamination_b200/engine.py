"""Host-side driver of the dense CLIP segmentation hot path on one B200.

Three engines, each a sequence of libclipseg kernel launches on the current CUDA stream:

* ``VisualEngine``  -- open_clip ViT dense patch-feature pass (open_clip/transformer.py:538-775)
* ``JBUEngine``     -- SimFeatUp JBUOne / JBUStack upsampler (simfeatup_dev/upsamplers.py:278-325)
* ``SegEngine``     -- forward_slide + forward_feature head + postprocess_result
                       (segmentor.py:286-392,394-451,475-499)

All crops of an image are batched into the M dimension of the GEMMs (the reference runs them one
at a time).  Activations feeding GEMMs are bf16 (``precision='bf16'``, tcgen05 path) or fp32
(``precision='fp32'``, CUDA-core verification mode); the residual stream, LayerNorm statistics,
softmax, similarity map and logits are fp32 in both modes.  PyTorch supplies device memory and
streams only -- there is no torch compute on the per-image path and no CPU fallback.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import ops
from ._lib import ATTN, ACT_NONE, ACT_GELU, ACT_QUICKGELU


def _round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def _on_device(fn):
    """Run an engine entry point with the engine's device current: kernels launch on the current device and
    ops._stream() takes its current stream, whatever device the caller had selected (SegmentorEx(device='cuda:1'))."""
    import functools

    @functools.wraps(fn)
    def inner(self, *a, **k):
        with torch.cuda.device(self.device):
            return fn(self, *a, **k)
    return inner


class Workspace:
    """Grow-only named device buffers (no allocation on the steady-state path).

    ``generation`` counts re-allocations: a CUDA graph captured while the buffers had generation g holds raw pointers
    into them and must not be replayed once any buffer has been replaced (SegEngine.graph checks this and recaptures)."""

    def __init__(self, device):
        self.device = device
        self._bufs: Dict[str, torch.Tensor] = {}
        self.generation = 0

    def get(self, name: str, shape: Sequence[int], dtype: torch.dtype, zero: bool = False) -> torch.Tensor:
        """zero: the buffer is zero-filled when it is (re)allocated (never on the steady-state path)."""
        n = 1
        for s in shape:
            n *= int(s)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        buf = self._bufs.get(name)
        if buf is None or buf.numel() < nbytes:
            if buf is not None:
                self.generation += 1              # an existing buffer is replaced: captured graphs are stale
            buf = (torch.zeros if zero else torch.empty)(max(nbytes, 256), dtype=torch.uint8, device=self.device)
            self._bufs[name] = buf
        return buf[:nbytes].view(dtype).view(*shape)

    def nbytes(self) -> int:
        return sum(b.numel() for b in self._bufs.values())


def slide_windows(h_img: int, w_img: int, stride: int, crop: int) -> List[Tuple[int, int, int, int]]:
    """forward_slide window enumeration, segmentor.py:411-423: (y1, x1, h, w), last window snapped back."""
    hg = max(h_img - crop + stride - 1, 0) // stride + 1
    wg = max(w_img - crop + stride - 1, 0) // stride + 1
    out = []
    for hi in range(hg):
        for wi in range(wg):
            y1, x1 = hi * stride, wi * stride
            y2, x2 = min(y1 + crop, h_img), min(x1 + crop, w_img)
            y1, x1 = max(y2 - crop, 0), max(x2 - crop, 0)
            out.append((y1, x1, y2 - y1, x2 - x1))
    return out


def compute_padsize(H: int, W: int, patch: int):
    """segmentor.py:534-546 -> (left, right, top, bottom)."""
    l = r = t = b = 0
    if W % patch:
        lr = patch - (W % patch)
        l = lr // 2
        r = lr - l
    if H % patch:
        tb = patch - (H % patch)
        t = tb // 2
        b = tb - t
    return l, r, t, b


class VisualEngine:
    """ViT image tower.  ``sd``: visual-tower weights keyed relative to ``visual.``
    (conv1.weight, class_embedding, positional_embedding, ln_pre.*, transformer.resblocks.N.*, ln_post.*,
    proj), any float dtype; repacked once to kernel layouts."""

    def __init__(self, sd: Dict[str, torch.Tensor], *, width: int, layers: int, heads: int, patch_size: int,
                 image_size: int, embed_dim: int, quick_gelu: bool = False, precision: str = 'bf16',
                 device='cuda'):
        assert precision in ('bf16', 'fp32')
        self.device = torch.device(device)
        self.cdt = torch.bfloat16 if precision == 'bf16' else torch.float32
        self.precision = precision
        self.width, self.layers, self.heads, self.ps = width, layers, heads, patch_size
        self.head_dim = width // heads
        self.grid0 = image_size // patch_size
        self.D = embed_dim
        self.act = ACT_QUICKGELU if quick_gelu else ACT_GELU
        f32 = lambda t: t.detach().to(self.device, torch.float32).contiguous()
        cd = lambda t: t.detach().to(self.device, torch.float32).to(self.cdt).contiguous()
        kk = 3 * patch_size * patch_size
        self.Kp = _round_up(kk, 64)
        wc = torch.zeros(width, self.Kp, dtype=torch.float32, device=self.device)
        wc[:, :kk] = f32(sd['conv1.weight']).reshape(width, kk)
        self.w_conv = wc.to(self.cdt).contiguous()
        self.cls_emb = f32(sd['class_embedding'])
        self.pos = f32(sd['positional_embedding'])
        self._pos_cache: Dict[Tuple[int, int], torch.Tensor] = {}
        self.ln_pre = (f32(sd['ln_pre.weight']), f32(sd['ln_pre.bias']))
        self.ln_post = (f32(sd['ln_post.weight']), f32(sd['ln_post.bias']))
        self.projT = cd(sd['proj'].t())                                  # [D, width] = B operand [N, K]
        self.blocks = []
        for i in range(layers):
            p = f'transformer.resblocks.{i}.'
            self.blocks.append(dict(
                ln1=(f32(sd[p + 'ln_1.weight']), f32(sd[p + 'ln_1.bias'])),
                ln2=(f32(sd[p + 'ln_2.weight']), f32(sd[p + 'ln_2.bias'])),
                w_in=cd(sd[p + 'attn.in_proj_weight']), b_in=f32(sd[p + 'attn.in_proj_bias']),
                w_out=cd(sd[p + 'attn.out_proj.weight']), b_out=f32(sd[p + 'attn.out_proj.bias']),
                w_fc=cd(sd[p + 'mlp.c_fc.weight']), b_fc=f32(sd[p + 'mlp.c_fc.bias']),
                w_pr=cd(sd[p + 'mlp.c_proj.weight']), b_pr=f32(sd[p + 'mlp.c_proj.bias'])))
        self.mlp = self.blocks[0]['w_fc'].shape[0]
        self.ws = Workspace(self.device)

    def _pos_for(self, gh: int, gw: int, crop_h: int, crop_w: int) -> torch.Tensor:
        """positional embedding for a gh x gw grid; interpolate_pos_encoding
        (open_clip/transformer.py:777-795) when the token count differs from the pretraining grid.
        One-time parameter transform per grid shape (cached), not on the per-image path."""
        if gh * gw + 1 == self.pos.shape[0]:
            return self.pos
        key = (gh, gw)
        if key not in self._pos_cache:
            N = self.pos.shape[0] - 1
            g = int(math.sqrt(N))
            w0, h0 = crop_h // self.ps + 0.1, crop_w // self.ps + 0.1
            pp = F.interpolate(self.pos[1:].reshape(1, g, g, -1).permute(0, 3, 1, 2),
                               scale_factor=(w0 / math.sqrt(N), h0 / math.sqrt(N)), mode='bicubic')
            assert int(w0) == pp.shape[-2] and int(h0) == pp.shape[-1]
            pp = pp.permute(0, 2, 3, 1).reshape(-1, self.width)
            self._pos_cache[key] = torch.cat([self.pos[:1], pp], 0).contiguous()
        return self._pos_cache[key]

    @_on_device
    def encode(self, img: torch.Tensor, windows: torch.Tensor, crop_h: int, crop_w: int, pad_top: int = 0,
               pad_left: int = 0, model_type: str = 'Experimental', ignore_residual: bool = True,
               sim_cfg: Optional[dict] = None, outlier_cfg: Optional[dict] = None,
               taps: Optional[dict] = None) -> Tuple[torch.Tensor, int]:
        """img: ops.Image (or a normalised fp32 [3,H,W] tensor), windows int32 [n,4] (y1,x1,h,w) in canvas rows.
        Returns (tok fp32 [n*L, D] = ln_post(.) @ proj for CLS + patches, L)."""
        if model_type not in ATTN or model_type == 'STD':
            raise NotImplementedError(f'model_type {model_type!r} is not supported by the CUDA path')
        ws, n, width, cdt = self.ws, windows.shape[0], self.width, self.cdt
        ps = self.ps
        gh, gw = crop_h // ps, crop_w // ps
        P = gh * gw
        L = P + 1
        M = n * L
        f32 = torch.float32
        patches = ws.get('patches', (n * P, self.Kp), cdt)
        ops.patchify(img, windows, crop_h, crop_w, pad_top, pad_left, ps, patches)
        pe = ws.get('pe', (n * P, width), f32)
        ops.gemm(patches, self.w_conv, pe)
        x = ws.get('x', (M, width), f32)
        if width % 4 == 0 and width <= 1280:
            ops.embed_tokens_ln(pe, self.cls_emb, self._pos_for(gh, gw, crop_h, crop_w), n, L, width, *self.ln_pre, x)
        else:
            ops.embed_tokens(pe, self.cls_emb, self._pos_for(gh, gw, crop_h, crop_w), n, L, width, x)
            ops.layernorm(x, *self.ln_pre, out=x)
        if taps is not None:
            taps['ln_pre'] = x.clone()
        h = ws.get('h', (M, width), cdt)
        qkv = ws.get('qkv', (M, 3 * width), cdt)
        att = ws.get('att', (M, width), cdt)
        g = ws.get('mlp', (M, self.mlp), cdt)
        use_sim = sim_cfg is not None
        use_out = outlier_cfg is not None
        mid_idx = (self.layers - 1) // 2                                   # transformer.py:593
        # bf16, model_type 'Experimental', head_dim 64, L <= 272: the final block runs on tcgen05 and reads the similarity map
        # in its padded, row-block-transposed layout (ops.simmap(transposed=True)); taps keep the plain [n, P, P] map
        exp_tc = (cdt == torch.bfloat16 and model_type == 'Experimental' and self.head_dim == 64 and 17 <= L <= ops.SIMT_COLS_MAX
                  and taps is None and os.environ.get('CSEG_ATTN_TC', '1') != '0')
        sim_w = (sim_cfg or {}).get('similarity_weight', 1.0)
        sim_t = (use_sim and exp_tc and sim_cfg.get('add_self_similarity', True) and width % 64 == 0 and width <= 1280)
        if sim_t:
            # one buffer per layout (208 / 272 key columns): only entries i, j >= 1 are ever written, so the zero CLS row /
            # column of a layout survive any sequence of L and batch sizes -- but not a change of layout
            simmap = ws.get(f'simmap_t{ops.simt_cols(L)}', (n, ops.simt_floats(L)), f32, zero=True)
        else:
            simmap = ws.get('simmap', (n, P, P), f32) if use_sim else None
        stats = ws.get('stats', (n, self.heads, 2, P), f32) if use_out else None
        have_sim = False
        for idx in range(self.layers - 1):
            b = self.blocks[idx]
            if idx == mid_idx and use_sim:
                keep_diag = sim_cfg.get('add_self_similarity', True)
                tc = cdt == torch.bfloat16 and keep_diag and width % 64 == 0 and width <= 1280
                ops.simmap(x, n, L, width, simmap, sim_cfg.get('temperature', 1.0), keep_diag,
                           scratch=ws.get('simmap_split', (M, 2 * width), cdt) if tc else None, transposed=sim_t)
                have_sim = True
            ops.layernorm(x, *b['ln1'], out=h)
            ops.gemm(h, b['w_in'], qkv, bias=b['b_in'])
            ops.attention(qkv, n, L, self.heads, self.head_dim, ATTN['STD'], att,
                          stats=stats if (use_out and idx == self.layers - 2) else None)
            ops.gemm(att, b['w_out'], x, bias=b['b_out'], residual=x)
            ops.layernorm(x, *b['ln2'], out=h)
            ops.gemm(h, b['w_fc'], g, bias=b['b_fc'], act=self.act)
            ops.gemm(g, b['w_pr'], x, bias=b['b_pr'], residual=x)
            if taps is not None:
                taps[f'block{idx}'] = x.clone()
        if taps is not None and have_sim:
            taps['simmap'] = simmap.clone()
        b = self.blocks[-1]
        ops.layernorm(x, *b['ln1'], out=h)
        ops.gemm(h, b['w_in'], qkv, bias=b['b_in'])
        if exp_tc and (sim_t or not have_sim):
            ops.attention_experimental_tc(qkv, n, L, self.heads, att, simmap if have_sim else None, sim_w)
        else:
            ops.attention(qkv, n, L, self.heads, self.head_dim, ATTN[model_type], att,
                          simmap=simmap if have_sim else None, sim_weight=sim_w)
        y = ws.get('y', (M, width), f32)
        if ignore_residual:                                                # transformer.py:627-628
            ops.gemm(att, b['w_out'], y, bias=b['b_out'])
        else:                                                              # transformer.py:641-643
            ops.gemm(att, b['w_out'], y, bias=b['b_out'], residual=x)
            ops.layernorm(y, *b['ln2'], out=h)
            ops.gemm(h, b['w_fc'], g, bias=b['b_fc'], act=self.act)
            ops.gemm(g, b['w_pr'], y, bias=b['b_pr'], residual=y)
        if taps is not None:
            taps['final_attn'] = y.clone()
            if use_out and self.layers >= 2:
                taps['stats'] = stats.clone()
        if use_out and self.layers >= 2:                                   # transformer.py:721-742
            if gh != gw:
                raise NotImplementedError('outlier suppression assumes square crops (transformer.py:583)')
            k = int(outlier_cfg.get('top_k', 10))
            plan = ws.get('outlier_plan', (n * (25 * k + P),), torch.int32)
            oidx = ws.get('outlier_idx', (n, k), torch.int32)
            y2 = ws.get('y2', (M, width), f32)
            ops.outlier_suppress(y, y2, n, L, width, gh, stats, self.heads, k,
                                 float(outlier_cfg.get('contamination_temp', 0.1)), plan, oidx)
            y = y2
            if taps is not None:
                taps['outlier_idx'] = oidx.clone()
                taps['suppressed'] = y.clone()
        ops.layernorm(y, *self.ln_post, out=h)
        tok = ws.get('tok', (M, self.D), f32)
        ops.gemm(h, self.projT, tok)
        return tok, L


class TextEngine:
    """CLIP text tower (open_clip/model.py:288-306, transformer.py:1047-1053) on the same kernels as the ViT: token +
    positional embedding gather, 12 pre-LN blocks (tcgen05 GEMMs, causal tensor-core attention), ln_final on the EOT
    rows, text_projection.  Init-time only (the prompt-ensembled class embeddings are cached by the segmentor).
    ``sd``: text-side weights keyed as in the CLIP state dict (token_embedding.weight, positional_embedding,
    transformer.resblocks.N.*, ln_final.*, text_projection)."""

    def __init__(self, sd: Dict[str, torch.Tensor], *, width: int, layers: int, heads: int, quick_gelu: bool = False,
                 precision: str = 'bf16', device='cuda'):
        assert precision in ('bf16', 'fp32')
        self.device = torch.device(device)
        self.cdt = torch.bfloat16 if precision == 'bf16' else torch.float32
        self.width, self.layers, self.heads = width, layers, heads
        self.head_dim = width // heads
        self.act = ACT_QUICKGELU if quick_gelu else ACT_GELU
        f32 = lambda t: t.detach().to(self.device, torch.float32).contiguous()
        cd = lambda t: t.detach().to(self.device, torch.float32).to(self.cdt).contiguous()
        self.tok_emb = f32(sd['token_embedding.weight'])
        self.pos = f32(sd['positional_embedding'])
        self.L = self.pos.shape[0]
        self.ln_final = (f32(sd['ln_final.weight']), f32(sd['ln_final.bias']))
        self.projT = cd(sd['text_projection'].t())                        # [D, width] = B operand [N, K]
        self.D = self.projT.shape[0]
        self.blocks = []
        for i in range(layers):
            p = f'transformer.resblocks.{i}.'
            self.blocks.append(dict(
                ln1=(f32(sd[p + 'ln_1.weight']), f32(sd[p + 'ln_1.bias'])),
                ln2=(f32(sd[p + 'ln_2.weight']), f32(sd[p + 'ln_2.bias'])),
                w_in=cd(sd[p + 'attn.in_proj_weight']), b_in=f32(sd[p + 'attn.in_proj_bias']),
                w_out=cd(sd[p + 'attn.out_proj.weight']), b_out=f32(sd[p + 'attn.out_proj.bias']),
                w_fc=cd(sd[p + 'mlp.c_fc.weight']), b_fc=f32(sd[p + 'mlp.c_fc.bias']),
                w_pr=cd(sd[p + 'mlp.c_proj.weight']), b_pr=f32(sd[p + 'mlp.c_proj.bias'])))
        self.mlp = self.blocks[0]['w_fc'].shape[0]
        self.ws = Workspace(self.device)

    def encode(self, tokens: torch.Tensor) -> torch.Tensor:
        """tokens int64 [n, context_length] -> fp32 [n, D] (un-normalised, like CLIP.encode_text)."""
        with torch.cuda.device(self.device):
            tokens = tokens.to(self.device, torch.int64).contiguous()
            n, L = tokens.shape
            assert L == self.L
            ws, width, cdt, f32 = self.ws, self.width, self.cdt, torch.float32
            M = n * L
            x = ws.get('x', (M, width), f32)
            ops.gather_rows(self.tok_emb, tokens.view(-1), self.pos, L, x)              # model.py:291-293
            h = ws.get('h', (M, width), cdt)
            qkv = ws.get('qkv', (M, 3 * width), cdt)
            att = ws.get('att', (M, width), cdt)
            g = ws.get('mlp', (M, self.mlp), cdt)
            for b in self.blocks:                                                       # transformer.py:234-254
                ops.layernorm(x, *b['ln1'], out=h)
                ops.gemm(h, b['w_in'], qkv, bias=b['b_in'])
                ops.attention(qkv, n, L, self.heads, self.head_dim, ATTN['CAUSAL'], att)
                ops.gemm(att, b['w_out'], x, bias=b['b_out'], residual=x)
                ops.layernorm(x, *b['ln2'], out=h)
                ops.gemm(h, b['w_fc'], g, bias=b['b_fc'], act=self.act)
                ops.gemm(g, b['w_pr'], x, bias=b['b_pr'], residual=x)
            # ln_final is row-wise, so it is applied to the n EOT rows only (EOT has the highest id: argmax, model.py:302-304)
            rows = torch.arange(n, device=self.device, dtype=torch.int64) * L + tokens.argmax(dim=-1)
            Mp = _round_up(n, 8)
            xe = ws.get('xe', (Mp, width), f32)
            xe[n:].zero_()
            ops.gather_rows(x, rows, None, 1, xe[:n])
            he = ws.get('he', (Mp, width), cdt)
            ops.layernorm(xe, *self.ln_final, out=he)
            out = torch.empty((Mp, self.D), dtype=f32, device=self.device)
            ops.gemm(he, self.projT, out)
            return out[:n]


class JBUEngine:
    """JBUOne ('up.*', radius 5, one module reused 4x) / JBUStack ('up1..4.*', radius 3)."""

    def __init__(self, name: str, sd: Dict[str, torch.Tensor], feat_dim: int, precision: str = 'bf16',
                 device='cuda'):
        if name == 'jbu_one':
            mods = [('up.', 5)] * 4
        elif name == 'jbu_stack':
            mods = [(f'up{i}.', 3) for i in range(1, 5)]
        else:
            raise ValueError(f"Unknown upsampler {name}")
        self.name, self.C = name, feat_dim
        self.device = torch.device(device)
        self.cdt = torch.bfloat16 if precision == 'bf16' else torch.float32
        f32 = lambda t: t.detach().to(self.device, torch.float32).contiguous()
        self.stages = []
        cache = {}
        for pre, r in mods:
            if pre not in cache:
                d2 = (2 * r + 1) ** 2
                ldk = _round_up(d2 + 3, 64)
                w0 = torch.zeros(ldk, ldk, device=self.device)
                w0[:d2, :d2 + 3] = f32(sd[pre + 'fixup_proj.0.weight']).reshape(d2, d2 + 3)
                b0 = torch.zeros(ldk, device=self.device)
                b0[:d2] = f32(sd[pre + 'fixup_proj.0.bias'])
                # second fix-up conv as ONE GEMM without a residual read: kernel += 0.1 * (W3 . hidden + b3) is
                # [hidden | kernel] . [0.1 W3 | I]^T + 0.1 b3  (the identity block passes the bf16 kernel exactly)
                w3 = torch.zeros(ldk, 2 * ldk, device=self.device)
                w3[:d2, :d2] = 0.1 * f32(sd[pre + 'fixup_proj.3.weight']).reshape(d2, d2)
                w3[:, ldk:] = torch.eye(ldk, device=self.device)
                b3 = torch.zeros(ldk, device=self.device)
                b3[:d2] = 0.1 * f32(sd[pre + 'fixup_proj.3.bias'])
                cache[pre] = dict(
                    radius=r, ldk=ldk,
                    range_temp=float(sd[pre + 'range_temp']), sigma=float(sd[pre + 'sigma_spatial']),
                    rp_w0=f32(sd[pre + 'range_proj.0.weight']).reshape(32, 3).contiguous(),
                    rp_b0=f32(sd[pre + 'range_proj.0.bias']),
                    rp_w3=f32(sd[pre + 'range_proj.3.weight']).reshape(32, 32).contiguous(),
                    rp_b3=f32(sd[pre + 'range_proj.3.bias']),
                    fx_w0=w0.to(self.cdt).contiguous(), fx_b0=b0, fx_w3=w3.to(self.cdt).contiguous(), fx_b3=b3)
            self.stages.append(cache[pre])
        self.w_fin = f32(sd['fixup_proj.1.weight']).reshape(feat_dim, feat_dim).to(self.cdt).contiguous()
        self.b_fin = f32(sd['fixup_proj.1.bias'])
        self.ws = Workspace(self.device)

    # ---- kernels shared across overlapping crops (csrc/jbu_share.cuh) ----------------------------------------------
    SHARED_STAGES = (2, 3)        # the 112^2 and 224^2 stages: 94 % of the pixels; the frames of stages 0/1 are most of the crop

    def share_ok(self, wl, n_canvas_px: int, crop_h: int, crop_w: int, pad_top: int, pad_left: int) -> bool:
        """Guidance pooling -> range projection -> range kernel -> fix-up -> composite kernels depend on the image, not
        on the crop, away from the crop border.  Sharing needs full-size windows on a 16-pixel lattice (pooling windows
        and bicubic phases then coincide between crops) and pays when the windows overlap (>= 2x coverage)."""
        if self.cdt != torch.bfloat16 or os.environ.get('CSEG_JBU_SHARE', '1') == '0':
            return False
        if pad_top or pad_left or crop_h % 16 or crop_w % 16 or crop_h < 16 * 4 or crop_w < 16 * 4:
            return False
        if any(st['ldk'] not in (64, 128) for st in self.stages):
            return False
        if any((y1 % 16) or (x1 % 16) or (h, w) != (crop_h, crop_w) for (y1, x1, h, w) in wl):
            return False
        return len(wl) * crop_h * crop_w >= 2 * n_canvas_px

    @_on_device
    def prepare_shared(self, img, crop_h: int, crop_w: int, gh: int, gw: int) -> dict:
        """Image-level tensors of the shared stages for the whole canvas of `img` (one region = all stacked images; the
        seams between images lie inside every crop's border frame, which is computed per crop)."""
        ws = self.ws
        img = ops._image(img)
        Hc, W = img.H, img.W
        key = (Hc, W)
        if getattr(self, '_full_win', None) is None:
            self._full_win = {}
        if key not in self._full_win:
            self._full_win[key] = torch.tensor([(0, 0, Hc, W)], dtype=torch.int32, device=self.device)
        full = self._full_win[key]
        out = {}
        for si in self.SHARED_STAGES:
            st = self.stages[si]
            GH, GW = gh << (si + 1), gw << (si + 1)                      # crop region at this stage
            shift = (crop_h // GH).bit_length() - 1                      # image pixels per stage pixel = 2^shift
            IH, IW = Hc >> shift, W >> shift
            npx, kw = IH * IW, st['ldk']
            guid = ws.get(f'img_guid{si}', (npx, 4), torch.float32)
            proj = ws.get(f'img_proj{si}', (npx, 32), torch.float16)
            ops.jbu_guidance_proj(img, full, Hc, W, 0, 0, IH, IW, st['rp_w0'], st['rp_b0'], st['rp_w3'], st['rp_b3'], guid, proj)
            kraw = ws.get(f'img_kraw{si}', (npx, kw), self.cdt)
            kern = ws.get(f'img_kern{si}', (npx, kw), self.cdt)
            ops.jbu_range_kernel(proj, guid, 1, IH, IW, st['radius'], st['range_temp'], st['sigma'], kraw)
            ops.jbu_kernel_fixup(kraw, st['fx_w0'], st['fx_b0'], st['fx_w3'][:, :kw], st['fx_b3'], kern)
            kc = ws.get(f'img_kc{si}', (npx, 128), self.cdt)
            tabs = ws.get('img_tabs', ((GH + GW) * 512,), torch.uint8)
            ops.jbu_composite_image(kern, IH, IW, GH, GW, st['radius'], kc, tabs)
            out[si] = dict(guid=guid, proj=proj, kern=kern, kc=kc, shift=shift, pitch=IW)
        return out

    @_on_device
    def upsample(self, feats: torch.Tensor, gh: int, gw: int, img: torch.Tensor, windows: torch.Tensor,
                 crop_h: int, crop_w: int, pad_top: int = 0, pad_left: int = 0,
                 taps: Optional[dict] = None, final_conv: bool = True, shared: Optional[dict] = None) -> torch.Tensor:
        """feats T [n*gh*gw, C] channel-last; returns T [n*(16gh)*(16gw), C] after the final fix-up
        (upsamplers.py:320-325); with final_conv=False the output of the fourth stage (the caller fuses the
        1x1 conv with the normalise + similarity, ``ops.fixup_norm_sim``)."""
        ws, n, C, cdt = self.ws, windows.shape[0], feats.shape[1], self.cdt     # C = self.C, or the basis width
        f32 = torch.float32
        s, h, w = feats, gh, gw
        for si, st in enumerate(self.stages):
            GH, GW = 2 * h, 2 * w
            npix = n * GH * GW
            if shared is not None and si in shared:
                # border frames per crop from the image-level projections; interior pixels read the image-level tensors
                sh, kw = shared[si], st['ldk']
                rb = ops.jbu_share_rows(GH, GW, ops.FB_RANGE)
                kraw_b = ws.get('kraw_b', (n * rb, kw), cdt)
                kern_b = ws.get('kern_b', (n * rb, kw), cdt)
                ops.jbu_range_kernel_border(sh['proj'], sh['guid'], windows, sh['shift'], sh['pitch'], n, GH, GW,
                                            st['radius'], st['range_temp'], st['sigma'], kraw_b)
                ops.jbu_kernel_fixup(kraw_b, st['fx_w0'], st['fx_b0'], st['fx_w3'][:, :kw], st['fx_b3'], kern_b)
                hr = ws.get('hr', (npix, C), cdt)
                dst = ws.get(f'up{si % 2}', (npix, C), cdt)
                ops.jbu_apply_shared(s, n, h, w, C, kern_b, sh['kern'], sh['kc'], windows, sh['shift'], sh['pitch'],
                                     st['radius'], dst, hr)
                s, h, w = dst, GH, GW
                continue
            guid = ws.get('guid', (npix, 4), f32)
            proj = ws.get('proj', (npix, 32), torch.float16 if cdt == torch.bfloat16 else f32)
            if cdt == torch.bfloat16 and GW >= 16:    # pooling + MLP in one kernel
                ops.jbu_guidance_proj(img, windows, crop_h, crop_w, pad_top, pad_left, GH, GW, st['rp_w0'], st['rp_b0'],
                                      st['rp_w3'], st['rp_b3'], guid, proj)
            else:
                ops.jbu_guidance(img, windows, crop_h, crop_w, pad_top, pad_left, GH, GW, guid)
                ops.jbu_range_proj(guid, npix, st['rp_w0'], st['rp_b0'], st['rp_w3'], st['rp_b3'], proj)
            kw = st['ldk']
            kern = ws.get('kern', (npix, kw), cdt)
            if cdt == torch.bfloat16 and kw in (64, 128):
                # both fix-up convs in one kernel: hidden activations stay on chip
                kraw = ws.get('kraw', (npix, kw), cdt)
                ops.jbu_range_kernel(proj, guid, n, GH, GW, st['radius'], st['range_temp'], st['sigma'], kraw)
                ops.jbu_kernel_fixup(kraw, st['fx_w0'], st['fx_b0'], st['fx_w3'][:, :kw], st['fx_b3'], kern)
            else:
                hk = ws.get('hidkern', (npix, 2 * kw), cdt)                            # [hidden | kernel] rows
                ops.jbu_range_kernel(proj, guid, n, GH, GW, st['radius'], st['range_temp'], st['sigma'], hk[:, kw:])
                ops.gemm(hk[:, kw:], st['fx_w0'], hk[:, :kw], bias=st['fx_b0'], act=ACT_GELU)   # fixup_proj.0 + GELU
                ops.gemm(hk, st['fx_w3'], kern, bias=st['fx_b3'])                      # kernel + .1 * fixup_proj.3
            if taps is not None:
                taps.setdefault('jbu_kernels', []).append(kern.clone())
            hr = ws.get('hr', (npix, C), cdt)
            dst = ws.get(f'up{si % 2}', (npix, C), cdt)
            ops.jbu_apply(s, n, h, w, C, kern, st['radius'], dst, hr)
            if taps is not None:
                taps.setdefault('jbu_stages', []).append(dst.clone())
            s, h, w = dst, GH, GW
        if not final_conv:
            return s
        out = ws.get('fin', (n * h * w, C), cdt)
        ops.gemm(s, self.w_fin, out, bias=self.b_fin, residual=s, alpha=0.1)
        return out

    # ---- basis form: upsample the identity instead of the features (cseg_basis_logits) ----------------
    def basis_ok(self, P: int, hw: int, Q: int) -> bool:
        """The stack is linear in its source and treats all channels alike, so for P tokens per crop < C it is
        cheaper to upsample one-hot token indicators (width round_up(P, 64)) and contract with the P x P Gram
        matrix of the token features afterwards.  bf16 only; fp32 stays on the literal path."""
        n2 = 16 if Q + 1 <= 16 else 32
        return (self.cdt == torch.bfloat16 and _round_up(P, 16) + n2 <= 256 and max(128, _round_up(P, 64)) < self.C
                and hw % 128 == 0 and Q <= 31 and self.C % 8 == 0)

    def _basis_state(self, n: int, P: int, text: torch.Tensor) -> dict:
        key = (n, P, text.data_ptr())
        st = getattr(self, '_basis', {}).get(key)
        if st is None:
            dev, C, Q = self.device, self.C, text.shape[0]
            Cb, Tp = max(128, _round_up(P, 64)), _round_up(P, 8)
            eye = torch.zeros(n, P, Cb, device=dev, dtype=self.cdt)
            eye[:, torch.arange(P), torch.arange(P)] = 1
            ldg = _round_up(n * Tp + 8, 8)
            b = 0.1 * self.b_fin                                   # bias of the final fix-up, upsamplers.py:325
            n2 = 16 if Q + 1 <= 16 else 32
            tb = torch.zeros(n2, C, device=dev, dtype=torch.float32)
            tb[:Q] = text
            tb[Q] = b
            consts = torch.cat([text @ b, (b @ b).reshape(1)]).contiguous()
            st = dict(Cb=Cb, Tp=Tp, ldg=ldg, eye=eye.reshape(n * P, Cb), tb=tb.to(self.cdt).contiguous(), consts=consts,
                      g=torch.zeros(n * Tp, C, device=dev, dtype=self.cdt),
                      gram=torch.zeros(n * Tp, ldg, device=dev, dtype=self.cdt),
                      aux=torch.zeros(n2, ldg, device=dev, dtype=self.cdt))
            if not hasattr(self, '_basis'):
                self._basis = {}
            self._basis[key] = st
        return st

    @_on_device
    def basis_logits(self, feats: torch.Tensor, gh: int, gw: int, img: torch.Tensor, windows: torch.Tensor,
                     crop_h: int, crop_w: int, pad_top: int, pad_left: int, text: torch.Tensor,
                     logits: torch.Tensor, cls_bias: Optional[torch.Tensor] = None,
                     shared: Optional[dict] = None) -> torch.Tensor:
        """Same result as upsample(final_conv=False) + ops.fixup_norm_sim: logits fp32 [n, Q, crop_h*crop_w].
        feats: [n * round_up(P, 8), C], every crop's P tokens followed by zero rows (ops.cls_debias rows_per_crop):
        the per-crop column blocks of gram / aux then start on 16-byte boundaries, which TMA requires."""
        n, P, Q = windows.shape[0], gh * gw, text.shape[0]
        st = self._basis_state(n, P, text)
        g, gram, aux, Tp = st['g'], st['gram'], st['aux'], st['Tp']
        assert feats.shape[0] == n * Tp
        # token features after the final 1x1 conv (without its bias): g = x + 0.1 * x . W^T
        ops.gemm(feats, self.w_fin, g, residual=feats, alpha=0.1)
        ops.gemm_blockdiag(g, g, gram[:, :n * Tp], Tp)     # per-crop Gram matrices: one launch, diagonal tiles only
        ops.gemm(st['tb'], g, aux[:, :n * Tp])                      # <g, text[q]> and <g, b>
        s = self.upsample(st['eye'], gh, gw, img, windows, crop_h, crop_w, pad_top, pad_left, None, final_conv=False,
                          shared=shared)
        return ops.basis_logits(s, st['Cb'], n, crop_h * crop_w, P, Tp, gram, aux, st['consts'], Q, logits, cls_bias)


class SegEngine:
    """forward_slide + forward_feature head + postprocess_result for one image at a time."""

    def __init__(self, visual: VisualEngine, query_features: torch.Tensor, query_idx: Sequence[int], *,
                 model_type: str = 'Experimental', ignore_residual: bool = True, prob_thd: float = 0.0,
                 logit_scale: float = 50.0, slide_stride: int = 112, slide_crop: int = 224,
                 cls_token_lambda: float = 0.0, global_debias_factor: float = 0.0, bg_idx: int = 0,
                 upsampler: Optional[JBUEngine] = None, sim_cfg: Optional[dict] = None,
                 outlier_cfg: Optional[dict] = None, jbu_chunk: int = 192, basis: bool = True):
        self.v = visual
        self.device = visual.device
        self.text = query_features.detach().to(self.device, torch.float32).contiguous()
        self.query_idx_list = [int(i) for i in query_idx]
        self.query_idx = torch.tensor(self.query_idx_list, dtype=torch.int32, device=self.device)
        self.Q = len(self.query_idx_list)
        self.K = max(self.query_idx_list) + 1
        self.model_type, self.ignore_residual = model_type, ignore_residual
        self.prob_thd, self.logit_scale = float(prob_thd), float(logit_scale)
        self.stride, self.crop = slide_stride, slide_crop
        self.cls_token_lambda, self.debias, self.bg_idx = float(cls_token_lambda), float(global_debias_factor), int(bg_idx)
        self.up = upsampler
        self.sim_cfg, self.outlier_cfg = sim_cfg, outlier_cfg
        self.jbu_chunk = int(os.environ.get('CSEG_JBU_CHUNK', jbu_chunk))     # crops per JBU pass (working set ~95 MB per crop)
        self.basis = basis          # bf16: upsample token indicators instead of features when that is cheaper
        self.share_kernels = True   # bf16: JBU kernel generation once per image pixel + per-crop border frames
        self.overlap_image_level = os.environ.get('CSEG_OVERLAP', '1') != '0'   # ... on a side stream next to the ViT
        self._side = None
        self.ws = Workspace(self.device)
        self._win_cache: Dict[tuple, Tuple[torch.Tensor, list]] = {}
        self._graphs: Dict[tuple, dict] = {}
        self.max_graphs = 8
        self.mean = [122.771, 116.746, 104.094]          # segmentor.py:64-67 (RGB)
        self.std = [68.501, 66.632, 70.323]

    def _windows(self, H: int, W: int, B: int = 1):
        """Windows of B equally sized images stacked vertically into one canvas of B*H rows: the batch dimension of
        slide_inference (segmentor.py:413-449 crops all B images per window) becomes B times as many crops, and no
        window straddles two images."""
        key = (H, W, B, self.stride, self.crop)
        if key not in self._win_cache:
            if self.crop > 0:
                wl = slide_windows(H, W, self.stride, self.crop)
            else:                                                   # whole-image path, segmentor.py:470-471
                wl = [(0, 0, H, W)]
            wl = [(y1 + b * H, x1, h, w) for b in range(B) for (y1, x1, h, w) in wl]
            self._win_cache[key] = (torch.tensor(wl, dtype=torch.int32, device=self.device), wl)
        return self._win_cache[key]

    def _as_image(self, img, batch: int = 1) -> Tuple[ops.Image, int]:
        """(ops.Image, B).  A plain fp32 tensor [3, B*H, W] is the legacy stacked canvas (B given by `batch`)."""
        if isinstance(img, ops.Image):
            assert batch in (1, img.B)
            return img, img.B
        im = ops.Image.normalised(img)
        assert im.H % batch == 0
        return im, batch

    @_on_device
    def crop_logits(self, img, taps: Optional[dict] = None, batch: int = 1):
        """Per-crop cosine logits (forward_feature, segmentor.py:286-392) for every window of `img` (an ops.Image, or
        a normalised fp32 [3,H,W] tensor on the device; with batch = B > 1 the tensor holds B images stacked to
        [3, B*H, W]).  Returns (logits fp32 [n,Q,lh,lw], geometry)."""
        img, batch = self._as_image(img, batch)
        H, W = img.H, img.W
        win_dev, wl = self._windows(H // batch, W, batch)
        n = len(wl)
        wh, ww = wl[0][2], wl[0][3]
        ps = self.v.ps
        if self.crop > 0:
            pl, pr, pt, pb = compute_padsize(wh, ww, ps)
        else:
            pl = pr = pt = pb = 0
            wh, ww = (wh // ps) * ps, (ww // ps) * ps                # conv stride drops the remainder
            if self.up is not None and (wh, ww) != (wl[0][2], wl[0][3]):
                # the reference's view(1, C, H*W) of the upsampled grid fails here too (segmentor.py:372)
                raise ValueError(f'whole-image inference with the x16 upsampler needs H, W multiples of {ps}, '
                                 f'got {wl[0][2]}x{wl[0][3]}')
        crop_h, crop_w = wh + pt + pb, ww + pl + pr
        gh, gw = crop_h // ps, crop_w // ps
        P = gh * gw
        # The image-level JBU tensors (guidance, range projection / kernel, fix-up, composite kernels of the shared stages)
        # depend on the image only: they run on a side stream next to the ViT (SIMT / MUFU work in the gaps of the
        # tensor-core GEMMs) and are joined before the upsampler starts.  Inside a graph capture this is a parallel branch.
        shared, side = None, None
        if (self.up is not None and ps == 16 and self.share_kernels and taps is None and self.crop > 0
                and self.up.share_ok(wl, H * W, crop_h, crop_w, pt, pl)):
            if self.overlap_image_level:
                cur = torch.cuda.current_stream(self.device)
                if self._side is None:
                    self._side = torch.cuda.Stream(self.device)
                side = self._side
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    shared = self.up.prepare_shared(img, crop_h, crop_w, gh, gw)
            else:
                shared = self.up.prepare_shared(img, crop_h, crop_w, gh, gw)
        tok, L = self.v.encode(img, win_dev, crop_h, crop_w, pt, pl, self.model_type, self.ignore_residual,
                               self.sim_cfg, self.outlier_cfg, taps)
        if side is not None:
            torch.cuda.current_stream(self.device).wait_stream(side)
        D, cdt, ws = self.v.D, self.v.cdt, self.ws
        basis = (self.up is not None and self.basis and taps is None and ps == 16
                 and self.up.basis_ok(P, crop_h * crop_w, self.Q))
        Pp = _round_up(P, 8) if basis else P                       # basis form: zero rows pad every crop to 16 bytes
        feats = ws.get('feats', (n * Pp, D), cdt)
        cls_unit = ws.get('cls_unit', (n, D), torch.float32)
        ops.cls_debias(tok, n, L, D, self.debias, feats, cls_unit, rows_per_crop=Pp)
        if taps is not None:
            taps['tok'] = tok.clone()
            taps['patch_feats'] = feats.clone()
        cls_bias = None
        if self.cls_token_lambda != 0:                              # segmentor.py:311,378-379
            cls_bias = ws.get('cls_bias', (n, self.Q), torch.float32)
            ops.gemm(cls_unit, self.text, cls_bias, alpha=self.cls_token_lambda)
        if self.up is not None:
            if ps != 16:
                raise ValueError('JBU upsamples x16 and only matches patch size 16 (segmentor.py:372)')
            logits = ws.get('logits', (n, self.Q, crop_h, crop_w), torch.float32)
            for c0 in range(0, n, self.jbu_chunk):
                c1 = min(n, c0 + self.jbu_chunk)
                if basis:
                    self.up.basis_logits(feats[c0 * Pp:c1 * Pp], gh, gw, img, win_dev[c0:c1], crop_h, crop_w, pt, pl,
                                         self.text, logits[c0:c1], cls_bias[c0:c1] if cls_bias is not None else None,
                                         shared=shared)
                    continue
                y = self.up.upsample(feats[c0 * P:c1 * P], gh, gw, img, win_dev[c0:c1], crop_h, crop_w, pt, pl,
                                     taps if c0 == 0 else None, final_conv=False, shared=shared)
                fused = (cdt == torch.bfloat16 and D % 128 == 0 and D <= 512 and self.Q <= 16)
                scratch = None if fused else self.up.ws.get('fin', ((c1 - c0) * crop_h * crop_w, D), cdt)
                ops.fixup_norm_sim(y, self.up.w_fin, c1 - c0, crop_h * crop_w, D, self.up.b_fin, 0.1, self.text,
                                   logits[c0:c1], cls_bias[c0:c1] if cls_bias is not None else None, scratch)
        else:
            logits = ws.get('logits', (n, self.Q, gh, gw), torch.float32)
            ops.norm_sim(feats, D, n, P, D, self.text, logits, cls_bias)
        geom = dict(windows=win_dev, crop_h=crop_h, crop_w=crop_w, pad_top=pt, pad_left=pl, H=H, W=W, n=n)
        return logits, geom

    def segment(self, img, ori_shape: Optional[Tuple[int, int]] = None, *, labels=None,
                want_probs: bool = False, want_logits: bool = False, taps: Optional[dict] = None, batch: int = 1):
        """predict() for one image: labels uint8 [out_h,out_w] (+ probs [K,..] / averaged logits [Q,H,W]).
        `img` is an ops.Image or a normalised fp32 tensor.  batch = B > 1 (or an Image of B images): every crop-level
        kernel works on B times as many crops per launch and the outputs are the B results stacked vertically
        (ori_shape must be None)."""
        with torch.cuda.device(self.device):
            img, batch = self._as_image(img, batch)
            assert batch == 1 or ori_shape is None
            logits, g = self.crop_logits(img, taps, batch)
            H, W = g['H'], g['W']
            out_h, out_w = (H, W) if ori_shape is None else (int(ori_shape[0]), int(ori_shape[1]))
            if labels is None:
                labels = torch.empty((out_h, out_w), dtype=torch.uint8, device=self.device)
            probs = torch.empty((self.K, out_h, out_w), dtype=torch.float32, device=self.device) if want_probs else None
            avg = None
            if want_logits and self.crop <= 0:          # whole image: forward_feature's output at ori_shape
                avg = torch.empty((self.Q, out_h, out_w), dtype=torch.float32, device=self.device)
            elif want_logits and (out_h, out_w) == (H, W):
                avg = torch.empty((self.Q, H, W), dtype=torch.float32, device=self.device)
            if self.crop > 0:
                ops.accum_argmax(logits, g['windows'], g['crop_h'], g['crop_w'], g['pad_top'], g['pad_left'], H, W,
                                 out_h, out_w, self.query_idx, self.K, self.logit_scale, self.prob_thd, self.bg_idx,
                                 labels, probs, avg)
            else:
                # whole image (segmentor.py:470-471): forward_feature interpolates the [lh, lw] logits straight to
                # ori_shape (one bilinear resize, :388-391): a single window of the output size
                win = self._whole_window(out_h, out_w)
                ops.accum_argmax(logits, win, out_h, out_w, 0, 0, out_h, out_w, out_h, out_w, self.query_idx, self.K,
                                 self.logit_scale, self.prob_thd, self.bg_idx, labels, probs, avg)
        return labels, probs, avg

    def _whole_window(self, h: int, w: int) -> torch.Tensor:
        key = ('whole', h, w)
        if key not in self._win_cache:
            self._win_cache[key] = (torch.tensor([(0, 0, h, w)], dtype=torch.int32, device=self.device), None)
        return self._win_cache[key][0]

    # ---- CUDA-graph replay of the whole per-batch launch sequence ----------------------------------
    def _generations(self):
        return (self.ws.generation, self.v.ws.generation, self.up.ws.generation if self.up is not None else 0)

    def _static_input(self, kind: str, B: int, H: int, W: int) -> torch.Tensor:
        if kind == 'u8hwc':
            return torch.zeros((B, H, W, 3), dtype=torch.uint8, device=self.device)
        if kind == 'u8chw':
            return torch.zeros((B, 3, H, W), dtype=torch.uint8, device=self.device)
        if kind == 'f32':
            return torch.zeros((B, 3, H, W), dtype=torch.float32, device=self.device)
        raise ValueError(kind)

    def _image_of(self, kind: str, t: torch.Tensor) -> ops.Image:
        if kind == 'f32':
            return ops.Image.normalised(t)
        return ops.Image.u8(t, 'hwc' if kind == 'u8hwc' else 'chw', self.mean, self.std)

    def graph(self, H: int, W: int, B: int = 1, kind: str = 'u8hwc',
              ori_shape: Optional[Tuple[int, int]] = None) -> dict:
        """Capture the launch sequence for B images of H x W once (a few hundred launches) and replay it afterwards:
        the per-launch host cost (Python, ctypes, tensor-map encodes) is paid at capture time only.
        kind: 'u8hwc' uint8 [B,H,W,3] BGR (cv2 / predict_u8), 'u8chw' uint8 [B,3,H,W] BGR (mmengine test_step),
        'f32' normalised [B,3,H,W] (predict).  Static buffers: st['in'] and st['labels'] (uint8 [B,out_h,out_w]).
        A graph holds raw pointers into the grow-only workspaces: it is discarded and recaptured when any of them has
        been re-allocated since the capture (a larger problem ran in between)."""
        key = (H, W, B, kind, ori_shape, self.stride, self.crop)
        st = self._graphs.get(key)
        if st is not None and st['gen'] == self._generations():
            if next(reversed(self._graphs)) != key:      # keep the dict in least-recently-used order
                self._graphs[key] = self._graphs.pop(key)
            return st
        while st is None and len(self._graphs) >= self.max_graphs:      # datasets with many image sizes: bound the cache
            self._graphs.pop(next(iter(self._graphs)))
        with torch.cuda.device(self.device):
            if st is None:
                oh, ow = (H, W) if ori_shape is None else (int(ori_shape[0]), int(ori_shape[1]))
                st = {'in': self._static_input(kind, B, H, W),
                      'labels': torch.empty((B, oh, ow), dtype=torch.uint8, device=self.device)}
                st['image'] = self._image_of(kind, st['in'])
            oh, ow = st['labels'].shape[1:]

            def run():
                self.segment(st['image'], ori_shape, labels=st['labels'].view(B * oh, ow))

            cur = torch.cuda.current_stream(self.device)
            side = torch.cuda.Stream(self.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                run()                                    # warm-up: sizes every workspace before the capture
                run()
            cur.wait_stream(side)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                run()
            st['graph'] = g
            st['gen'] = self._generations()
        self._graphs[key] = st
        return st

    def segment_batch(self, x, kind: str, ori_shape: Optional[Tuple[int, int]] = None,
                      labels_out: Optional[torch.Tensor] = None, use_graph: bool = True,
                      copy_out: bool = True) -> torch.Tensor:
        """B equally sized images -> uint8 labels [B,out_h,out_w] on the device.  `x`: one tensor laid out as `kind`
        (see graph()) or a list of B per-image tensors (what a dataloader hands over), pinned host or device memory.
        The batch runs as ONE launch sequence over B times as many crops.  With copy_out=False the returned tensor is
        the graph's static output buffer, overwritten by the next call."""
        xs = list(x) if isinstance(x, (list, tuple)) else None
        first = xs[0] if xs is not None else x[0]
        B = len(xs) if xs is not None else x.shape[0]
        H, W = (first.shape[0], first.shape[1]) if kind == 'u8hwc' else (first.shape[1], first.shape[2])
        assert B == 1 or ori_shape is None
        with torch.cuda.device(self.device):
            if not use_graph:
                if xs is not None:
                    xd = torch.stack([t.to(self.device, non_blocking=True) for t in xs])
                else:
                    xd = x.to(self.device, non_blocking=True).contiguous()
                oh, ow = (H, W) if ori_shape is None else (int(ori_shape[0]), int(ori_shape[1]))
                if labels_out is None:
                    labels_out = torch.empty((B, oh, ow), dtype=torch.uint8, device=self.device)
                self.segment(self._image_of(kind, xd), ori_shape, labels=labels_out.view(B * oh, ow))
                return labels_out
            st = self.graph(H, W, B, kind, ori_shape)
            on_host = not (xs[0] if xs is not None else x).is_cuda
            if on_host:
                # double-buffered upload on a copy stream: the H2D of this batch overlaps the replay of the previous one
                # (the caller's pinned buffers must stay valid until the copy has run, as with any non_blocking copy)
                if 'stage' not in st:
                    st['stage'] = [torch.empty_like(st['in']) for _ in range(2)]
                    st['copy_stream'] = torch.cuda.Stream(self.device)
                    st['ready'] = [torch.cuda.Event() for _ in range(2)]
                    st['free'] = [torch.cuda.Event() for _ in range(2)]
                    st['k'] = 0
                k = st['k'] = st['k'] ^ 1
                cur = torch.cuda.current_stream(self.device)
                cs = st['copy_stream']
                cs.wait_event(st['free'][k])             # staging[k] was last read two calls ago
                with torch.cuda.stream(cs):
                    if xs is not None:
                        for i, t in enumerate(xs):
                            st['stage'][k][i].copy_(t, non_blocking=True)
                    else:
                        st['stage'][k].copy_(x, non_blocking=True)
                    st['ready'][k].record(cs)
                cur.wait_event(st['ready'][k])
                st['in'].copy_(st['stage'][k], non_blocking=True)
                st['free'][k].record(cur)
            elif xs is not None:                         # device inputs: D2D straight into the static input
                for i, t in enumerate(xs):
                    st['in'][i].copy_(t, non_blocking=True)
            else:
                st['in'].copy_(x, non_blocking=True)
            st['graph'].replay()
            if labels_out is not None:
                labels_out.copy_(st['labels'], non_blocking=True)
                return labels_out
            return st['labels'].clone() if copy_out else st['labels']

    def segment_u8(self, img_hwc_bgr_u8: torch.Tensor, labels_out: Optional[torch.Tensor] = None,
                   use_graph: bool = True, copy_out: bool = True) -> torch.Tensor:
        """uint8 HWC BGR image [H,W,3] -> uint8 labels [H,W] on the device, or a batch [B,H,W,3] -> [B,H,W]
        (host pinned or device input)."""
        batched = img_hwc_bgr_u8.dim() == 4
        x = img_hwc_bgr_u8 if batched else img_hwc_bgr_u8.unsqueeze(0)
        lo = labels_out if (labels_out is None or batched) else labels_out.unsqueeze(0)
        out = self.segment_batch(x, 'u8hwc', None, lo, use_graph, copy_out)
        return out if batched else out[0]
