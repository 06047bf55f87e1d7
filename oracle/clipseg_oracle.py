"""CPU oracle for the dense CLIP segmentation hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a plain torch-fp32 (CPU) restatement of the reference algorithm
(UserNameUnavailableIsUnavailable/CLIP-Decontamination, a pure-Python repo).  It is the
checker for the CUDA path; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product
package (``clip_decontamination_b200``) never imports anything from ``oracle/``.

Parity status: PINNED.  The reference holds no golden vectors of its own, so
the oracle is pinned against the reference's own Python imported in the build container
(``oracle/ref_harness.py`` + ``oracle/gen_golden.py``) and the resulting vectors are
committed under ``tests/golden/``.  ``tests/test_oracle_golden.py`` re-checks the oracle
against those vectors on every run.

Every function cites the reference file:line it follows (paths relative to the reference
repo root).  All tensors are torch.float32 on the CPU unless noted; weights come in as a
plain ``dict[str, Tensor]`` with the reference's state-dict key names.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------
# ViT image tower (open_clip/transformer.py)
# ----------------------------------------------------------------------------------------------

def _ln(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    """LayerNormFp32.forward, open_clip/transformer.py:17-23 (fp32 compute, cast back)."""
    return F.layer_norm(x.float(), (x.shape[-1],), w.float(), b.float(), eps).to(x.dtype)


def _act(x: Tensor, quick_gelu: bool) -> Tensor:
    """nn.GELU (erf) for json configs, QuickGELU x*sigmoid(1.702x) for *-quickgelu / openai
    (open_clip/transformer.py:35-38, open_clip/model.py:116)."""
    if quick_gelu:
        return x * torch.sigmoid(1.702 * x)
    return F.gelu(x)


def _split_heads(t: Tensor, heads: int) -> Tensor:
    """[B, L, d] -> [B, h, L, hd]; equals view(L, B*h, hd).transpose(0,1) on the LND tensor
    (open_clip/transformer.py:842-844)."""
    B, L, d = t.shape
    return t.view(B, L, heads, d // heads).permute(0, 2, 1, 3)


def std_attention(x_ln: Tensor, p: Dict[str, Tensor], pre: str, heads: int,
                  need_weights: bool = False):
    """nn.MultiheadAttention forward as used by ResidualAttentionBlock.attention
    (open_clip/transformer.py:204,218-232).  x_ln: [B, L, d].  Returns out [B, L, d] and
    (optionally) the head-averaged softmax weights [B, L, L]."""
    B, L, d = x_ln.shape
    hd = d // heads
    qkv = F.linear(x_ln, p[pre + 'attn.in_proj_weight'], p[pre + 'attn.in_proj_bias'])
    q, k, v = qkv.chunk(3, dim=-1)
    q, k, v = _split_heads(q, heads), _split_heads(k, heads), _split_heads(v, heads)
    s = (q * (hd ** -0.5)) @ k.transpose(-1, -2)
    w = s.softmax(dim=-1)
    o = (w @ v).permute(0, 2, 1, 3).reshape(B, L, d)
    o = F.linear(o, p[pre + 'attn.out_proj.weight'], p[pre + 'attn.out_proj.bias'])
    if need_weights:
        return o, w.mean(dim=1)
    return o, None


def res_block(x: Tensor, p: Dict[str, Tensor], pre: str, heads: int, quick_gelu: bool,
              need_weights: bool = False):
    """ResidualAttentionBlock.forward, open_clip/transformer.py:234-254 (pre-LN)."""
    a, w = std_attention(_ln(x, p[pre + 'ln_1.weight'], p[pre + 'ln_1.bias']), p, pre, heads,
                         need_weights)
    x = x + a
    h = _ln(x, p[pre + 'ln_2.weight'], p[pre + 'ln_2.bias'])
    h = F.linear(h, p[pre + 'mlp.c_fc.weight'], p[pre + 'mlp.c_fc.bias'])
    h = _act(h, quick_gelu)
    h = F.linear(h, p[pre + 'mlp.c_proj.weight'], p[pre + 'mlp.c_proj.bias'])
    return x + h, w


def similarity_map(mid_patches: Tensor, temperature: float = 1.0,
                   add_self_similarity: bool = True) -> Tensor:
    """SimilarityEnhancementModule.compute_similarity_map, similarity_enhancement.py:37-66.
    mid_patches [B, P, d] -> [B, P, P] fp32 cosine similarities."""
    f = F.normalize(mid_patches.float(), p=2, dim=-1)
    m = torch.bmm(f, f.transpose(1, 2)) / temperature
    if not add_self_similarity:
        eye = torch.eye(m.shape[1], dtype=m.dtype).unsqueeze(0)
        m = m * (1 - eye)
    return m


def _pad_simmap(sim: Tensor, heads: int, weight: float, dtype) -> Tensor:
    """SimilarityEnhancementModule.enhance_attention, similarity_enhancement.py:78-124:
    zero CLS row/col, same map for every head, cast to the attention dtype."""
    B, P, _ = sim.shape
    m = torch.zeros(B, P + 1, P + 1, dtype=sim.dtype)
    m[:, 1:, 1:] = sim
    return (weight * m.to(dtype)).unsqueeze(1)  # [B,1,L,L] broadcast over heads


def custom_attention(x_ln: Tensor, p: Dict[str, Tensor], pre: str, heads: int,
                     model_type: str, sim: Optional[Tensor], sim_weight: float = 1.0) -> Tensor:
    """VisionTransformer.custom_attn, open_clip/transformer.py:822-940.  x_ln [B, L, d] is
    blk.ln_1(x).  Output: out_proj(W v) [B, L, d] -- no residual, no FFN."""
    B, L, d = x_ln.shape
    hd = d // heads
    scale = hd ** -0.5
    qkv = F.linear(x_ln, p[pre + 'attn.in_proj_weight'], p[pre + 'attn.in_proj_bias'])
    q, k, v = qkv.chunk(3, dim=-1)
    q, k, v = _split_heads(q, heads), _split_heads(k, heads), _split_heads(v, heads)
    add = _pad_simmap(sim, heads, sim_weight, q.dtype) if sim is not None else None

    def enh(a):
        return a + add if add is not None else a

    if model_type == 'vanilla':                                   # :858-863
        w = enh(q @ k.transpose(-1, -2) * scale).softmax(-1)
    elif model_type == 'MaskCLIP':                                # :864-869
        w = torch.eye(L, dtype=q.dtype).expand(B, heads, L, L)
    elif model_type == 'SCLIP':                                   # :870-877
        w = enh(q @ q.transpose(-1, -2) * scale).softmax(-1) + \
            enh(k @ k.transpose(-1, -2) * scale).softmax(-1)
    elif model_type == 'SegEarth':                                # :878-887
        w = enh(q @ q.transpose(-1, -2) * scale).softmax(-1) + \
            enh(k @ k.transpose(-1, -2) * scale).softmax(-1) + \
            enh(v @ v.transpose(-1, -2) * scale).softmax(-1)
    elif model_type == 'SFP':                                     # :888-895
        w = enh(0.5 * (q @ q.transpose(-1, -2) * scale + k @ k.transpose(-1, -2) * scale)).softmax(-1)
    elif model_type == 'Experimental':                            # :896-902
        kk = k @ k.transpose(-1, -2) * scale
        qq = q @ q.transpose(-1, -2) * scale
        w = (kk + qq).softmax(-1)
        w = enh(w)
        w = w.softmax(-1)                                         # second softmax unconditional
    elif model_type == 'ClearCLIP':                               # :903-908
        w = enh(q @ q.transpose(-1, -2) * scale).softmax(-1)
    else:
        raise NotImplementedError(model_type)
    o = (w @ v).permute(0, 2, 1, 3).reshape(B, L, d)
    return F.linear(o, p[pre + 'attn.out_proj.weight'], p[pre + 'attn.out_proj.bias'])


def detect_outliers(attn_avg: Tensor, num_patches: int, top_k: int) -> Tensor:
    """detect_outliers_by_attention, outlier_suppression.py:15-61.  attn_avg [B, L, L]."""
    diag = torch.diagonal(attn_avg, dim1=1, dim2=2)[:, 1:1 + num_patches]
    cls_row = attn_avg[:, 0, 1:1 + num_patches]
    ratio = cls_row / (diag + 1e-8)
    k = min(top_k, num_patches)
    return torch.topk(ratio, k=k, largest=True, dim=1).indices


def outlier_mean_interpolation(fmap: Tensor, outlier_idx: Tensor, contamination_temp: float = 0.1
                               ) -> Tensor:
    """OutlierSuppressionModule.mean_interpolation, outlier_suppression.py:115-214.
    fmap [B, C, H, W]; outlier_idx [B, k] flat patch indices in top-k order."""
    B, C, H, W = fmap.shape
    res = fmap.clone()
    offs = [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1)]
    for b in range(B):
        idx = outlier_idx[b].tolist()
        coords = [(i // W, i % W) for i in idx]
        repl = []
        for (oy, ox) in coords:
            o = fmap[b, :, oy, ox]
            nb_coords = [(min(max(oy + dy, 0), H - 1), min(max(ox + dx, 0), W - 1)) for dy, dx in offs]
            nb = torch.stack([fmap[b, :, y, x] for y, x in nb_coords])          # [8, C] ORIGINAL map
            sim = (F.normalize(nb, p=2, dim=1) * F.normalize(o[None], p=2, dim=1)).sum(1)
            w = torch.clamp(1.0 - sim, min=0.0).softmax(0)
            repl.append((nb * w[:, None]).sum(0))
            strength = torch.clamp(sim * contamination_temp, 0, 1)
            clean = nb - o[None] * strength[:, None]
            for j, (y, x) in enumerate(nb_coords):                              # last writer wins
                if y != oy or x != ox:
                    res[b, :, y, x] = clean[j]
        for (oy, ox), r in zip(coords, repl):                                   # outliers last
            res[b, :, oy, ox] = r
    return res


def interpolate_pos_encoding(pos: Tensor, n_tokens: int, w: int, h: int, patch: int) -> Tensor:
    """VisionTransformer.interpolate_pos_encoding, open_clip/transformer.py:777-795."""
    npatch = n_tokens - 1
    N = pos.shape[0] - 1
    if npatch == N and w == h:
        return pos
    dim = pos.shape[-1]
    w0, h0 = w // patch + 0.1, h // patch + 0.1
    g = int(math.sqrt(N))
    pp = F.interpolate(pos[1:].reshape(1, g, g, dim).permute(0, 3, 1, 2),
                       scale_factor=(w0 / math.sqrt(N), h0 / math.sqrt(N)), mode='bicubic')
    assert int(w0) == pp.shape[-2] and int(h0) == pp.shape[-1]
    pp = pp.permute(0, 2, 3, 1).reshape(1, -1, dim)
    return torch.cat((pos[[0]].unsqueeze(0), pp), dim=1)


def vit_dense_forward(p: Dict[str, Tensor], img: Tensor, *, layers: int, heads: int, patch: int,
                      quick_gelu: bool = False, model_type: str = 'Experimental',
                      ignore_residual: bool = True,
                      sim_cfg: Optional[dict] = None, outlier_cfg: Optional[dict] = None,
                      taps: Optional[dict] = None) -> Tuple[Tensor, Tensor]:
    """VisionTransformer.forward, open_clip/transformer.py:538-775 with last_n_layers=1,
    output_cls_token=True, layer fusion / self-attn enhancement off (segmentor defaults).

    p: visual-tower weights with keys relative to ``visual.`` (e.g. 'conv1.weight',
    'transformer.resblocks.0.ln_1.weight').  img [B,3,H,W].  Returns (cls [B,D], tokens
    [B,P,D]).  ``taps`` (optional dict) receives named intermediates for stage-wise tests."""
    B, _, Hh, Ww = img.shape
    x = F.conv2d(img, p['conv1.weight'], stride=patch)                     # :560 (no bias)
    x = x.reshape(B, x.shape[1], -1).permute(0, 2, 1)
    x = torch.cat([p['class_embedding'].to(x.dtype).expand(B, 1, -1), x], dim=1)   # :565
    pos = p['positional_embedding']
    if x.shape[1] != pos.shape[0]:
        x = x + interpolate_pos_encoding(pos, x.shape[1], Hh, Ww, patch).to(x.dtype)
    else:
        x = x + pos.to(x.dtype)
    x = _ln(x, p['ln_pre.weight'], p['ln_pre.bias'])                       # :574
    if taps is not None:
        taps['ln_pre'] = x.clone()
    grid = int(math.sqrt(x.shape[1] - 1))
    mid_idx = (layers - 1) // 2                                            # :593
    mid = None
    attn_w = None
    for idx in range(layers - 1):                                          # :591-612
        pre = f'transformer.resblocks.{idx}.'
        if idx == mid_idx and sim_cfg is not None:
            mid = x.clone()
        need = (idx == layers - 2) and outlier_cfg is not None
        x, w = res_block(x, p, pre, heads, quick_gelu, need_weights=need)
        if need:
            attn_w = w
        if taps is not None:
            taps[f'block{idx}'] = x.clone()
    sim = None
    if sim_cfg is not None and mid is not None:                            # :615-619
        sim = similarity_map(mid[:, 1:], sim_cfg.get('temperature', 1.0),
                             sim_cfg.get('add_self_similarity', True))
        if taps is not None:
            taps['simmap'] = sim.clone()
    pre = f'transformer.resblocks.{layers - 1}.'
    h = _ln(x, p[pre + 'ln_1.weight'], p[pre + 'ln_1.bias'])
    ca = custom_attention(h, p, pre, heads, model_type, sim,
                          (sim_cfg or {}).get('similarity_weight', 1.0))   # :628
    if ignore_residual:
        out = ca
    else:                                                                  # :641-643
        xo = x + ca
        hh = _ln(xo, p[pre + 'ln_2.weight'], p[pre + 'ln_2.bias'])
        hh = F.linear(hh, p[pre + 'mlp.c_fc.weight'], p[pre + 'mlp.c_fc.bias'])
        hh = _act(hh, quick_gelu)
        out = xo + F.linear(hh, p[pre + 'mlp.c_proj.weight'], p[pre + 'mlp.c_proj.bias'])
    if taps is not None:
        taps['final_attn'] = out.clone()
        if attn_w is not None:
            taps['attn_stats_cls'] = attn_w[:, 0, 1:].clone()
            taps['attn_stats_diag'] = torch.diagonal(attn_w, dim1=1, dim2=2)[:, 1:].clone()
    if outlier_cfg is not None and attn_w is not None:                     # :721-742
        P = out.shape[1] - 1
        fmap = out[:, 1:].permute(0, 2, 1).reshape(B, -1, grid, grid)
        oi = detect_outliers(attn_w, P, outlier_cfg.get('top_k', 10))
        if taps is not None:
            taps['outlier_idx'] = oi.clone()
        fmap = outlier_mean_interpolation(fmap, oi, outlier_cfg.get('contamination_temp', 0.1))
        out = torch.cat([out[:, :1], fmap.reshape(B, -1, P).permute(0, 2, 1)], dim=1)
        if taps is not None:
            taps['suppressed'] = out.clone()
    out = _ln(out, p['ln_post.weight'], p['ln_post.bias'])                 # :765
    cls, tokens = out[:, 0] @ p['proj'], out[:, 1:] @ p['proj']            # :766-770
    return cls, tokens


# ----------------------------------------------------------------------------------------------
# SimFeatUp JBU upsampler (simfeatup_dev/upsamplers.py)
# ----------------------------------------------------------------------------------------------

def adaptive_conv(src_padded: Tensor, filt: Tensor) -> Tensor:
    """Semantics of featup AdaptiveConv as pinned by adaptive_conv_py_simple,
    simfeatup_dev/upsamplers.py:14-25:  out[b,c,y,x] = sum_{i,j} in[b,c,y+i,x+j] * f[b,y,x,i,j].
    Tap-loop form (no 12 GB unfold)."""
    B, C, H1, W1 = src_padded.shape
    _, H2, W2, d, _ = filt.shape
    out = torch.zeros(B, C, H2, W2, dtype=src_padded.dtype)
    for i in range(d):
        for j in range(d):
            out.addcmul_(src_padded[:, :, i:i + H2, j:j + W2], filt[:, None, :, :, i, j])
    return out


def jbu_learned_range(p: Dict[str, Tensor], pre: str, source: Tensor, guidance: Tensor, radius: int,
                      taps: Optional[dict] = None) -> Tensor:
    """JBULearnedRange.forward, simfeatup_dev/upsamplers.py:253-275 (eval: Dropout2d = id)."""
    GB, GC, GH, GW = guidance.shape
    d = 2 * radius + 1
    # get_spatial_kernel :240-251
    dist = torch.linspace(-1, 1, d)
    gx, gy = torch.meshgrid(dist, dist, indexing='ij')
    sigma = p[pre + 'sigma_spatial']
    spatial = torch.exp(-(gx.square() + gy.square()) / (2 * sigma ** 2)).reshape(1, d * d, 1, 1)
    # get_range_kernel :230-238
    proj = F.conv2d(guidance, p[pre + 'range_proj.0.weight'], p[pre + 'range_proj.0.bias'])
    proj = F.gelu(proj)
    proj = F.conv2d(proj, p[pre + 'range_proj.3.weight'], p[pre + 'range_proj.3.bias'])
    pp = F.pad(proj, [radius] * 4, mode='reflect')
    pos_temp = p[pre + 'range_temp'].exp().clamp_min(1e-4).clamp_max(1e4)
    logits = torch.stack([(pp[:, :, i:i + GH, j:j + GW] * proj).sum(1)
                          for i in range(d) for j in range(d)], dim=1)      # [B, d*d, GH, GW]
    rng = F.softmax(pos_temp * logits, dim=1)
    comb = rng * spatial
    comb = comb / comb.sum(1, keepdim=True).clamp(1e-7)                    # :261-262
    fx = F.conv2d(torch.cat([comb, guidance], dim=1),
                  p[pre + 'fixup_proj.0.weight'], p[pre + 'fixup_proj.0.bias'])
    fx = F.gelu(fx)
    fx = F.conv2d(fx, p[pre + 'fixup_proj.3.weight'], p[pre + 'fixup_proj.3.bias'])
    comb = comb + 0.1 * fx                                                 # :264
    if taps is not None:
        taps.setdefault('jbu_kernels', []).append(comb.clone())
    filt = comb.permute(0, 2, 3, 1).reshape(GB, GH, GW, d, d)
    hr = F.interpolate(source, size=(GH, GW), mode='bicubic', align_corners=False)   # :268
    hrp = F.pad(hr, [radius] * 4, mode='reflect')
    return adaptive_conv(hrp, filt)


def jbu_upsample(p: Dict[str, Tensor], name: str, source: Tensor, guidance: Tensor,
                 taps: Optional[dict] = None) -> Tensor:
    """JBUOne.forward (:304-325; one radius-5 module reused 4x) / JBUStack.forward (:278-301;
    four radius-3 modules).  p: upsampler weights ('up.*'/'up1..4.*', 'fixup_proj.1.*')."""
    if name == 'jbu_one':
        mods = [('up.', 5)] * 4
    elif name == 'jbu_stack':
        mods = [(f'up{i}.', 3) for i in range(1, 5)]
    else:
        raise ValueError(f"Unknown upsampler {name}")
    s = source
    for pre, r in mods:
        h, w = s.shape[-2:]
        g = F.adaptive_avg_pool2d(guidance, (h * 2, w * 2))               # :316
        s = jbu_learned_range(p, pre, s, g, r, taps)
        if taps is not None:
            taps.setdefault('jbu_stages', []).append(s.clone())
    fix = F.conv2d(s, p['fixup_proj.1.weight'], p['fixup_proj.1.bias'])
    return fix * 0.1 + s                                                   # :325


# ----------------------------------------------------------------------------------------------
# Segmentor (segmentor.py)
# ----------------------------------------------------------------------------------------------

def get_cls_idx(path: str):
    """segmentor.py:611-622 -- one class per line, ',' separates synonyms, only '\\n' stripped."""
    with open(path, 'r') as f:
        lines = f.readlines()
    names, idx = [], []
    for i, line in enumerate(lines):
        parts = line.split(',')
        names += parts
        idx += [i] * len(parts)
    return [n.replace('\n', '') for n in names], idx


def compute_padsize(H: int, W: int, patch: int):
    """segmentor.py:534-546."""
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


def slide_windows(h_img: int, w_img: int, stride: int, crop: int) -> List[Tuple[int, int, int, int]]:
    """Window enumeration of forward_slide, segmentor.py:411-423 (last window snapped back)."""
    hg = max(h_img - crop + stride - 1, 0) // stride + 1
    wg = max(w_img - crop + stride - 1, 0) // stride + 1
    out = []
    for hi in range(hg):
        for wi in range(wg):
            y1, x1 = hi * stride, wi * stride
            y2, x2 = min(y1 + crop, h_img), min(x1 + crop, w_img)
            y1, x1 = max(y2 - crop, 0), max(x2 - crop, 0)
            out.append((y1, y2, x1, x2))
    return out


class SegOracle:
    """Dtype-parametrised restatement of SegmentorEx.forward_feature / forward_slide /
    postprocess_result / predict (segmentor.py:286-392, 394-451, 475-499, 453-473) and of
    Segmentor (segearth_segmentor.py) for the options both share."""

    def __init__(self, visual: Dict[str, Tensor], query_features: Tensor, query_idx: Sequence[int], *,
                 layers: int, heads: int, patch: int, quick_gelu: bool = False,
                 model_type: str = 'Experimental', ignore_residual: bool = True,
                 prob_thd: float = 0.0, logit_scale: float = 50, slide_stride: int = 112,
                 slide_crop: int = 224, cls_token_lambda: float = 0.0,
                 global_debias_factor: float = 0.0, bg_idx: int = 0,
                 upsampler: Optional[Tuple[str, Dict[str, Tensor]]] = None,
                 sim_cfg: Optional[dict] = None, outlier_cfg: Optional[dict] = None):
        self.visual = visual
        self.query_features = query_features.float()
        self.query_idx = torch.tensor(list(query_idx), dtype=torch.int64)
        self.num_queries = len(query_idx)
        self.num_classes = int(max(query_idx)) + 1
        self.layers, self.heads, self.patch, self.quick_gelu = layers, heads, patch, quick_gelu
        self.model_type, self.ignore_residual = model_type, ignore_residual
        self.prob_thd, self.logit_scale = prob_thd, logit_scale
        self.slide_stride, self.slide_crop = slide_stride, slide_crop
        self.cls_token_lambda, self.global_debias_factor, self.bg_idx = \
            cls_token_lambda, global_debias_factor, bg_idx
        self.upsampler = upsampler
        self.sim_cfg, self.outlier_cfg = sim_cfg, outlier_cfg

    # segmentor.py:286-392
    def forward_feature(self, img: Tensor, logit_size=None, taps: Optional[dict] = None) -> Tensor:
        cls, feats = vit_dense_forward(self.visual, img, layers=self.layers, heads=self.heads,
                                       patch=self.patch, quick_gelu=self.quick_gelu,
                                       model_type=self.model_type,
                                       ignore_residual=self.ignore_residual,
                                       sim_cfg=self.sim_cfg, outlier_cfg=self.outlier_cfg, taps=taps)
        B = img.shape[0]
        cls = cls / cls.norm(dim=-1, keepdim=True)                                   # :310
        cls_logits = cls @ self.query_features.T
        fw, fh = img.shape[-2] // self.patch, img.shape[-1] // self.patch            # :314
        if self.global_debias_factor != 0:                                           # :322-336
            fn = feats / feats.norm(dim=-1, keepdim=True)
            cn = cls / cls.norm(dim=-1, keepdim=True)
            s = (fn * cn.unsqueeze(1)).sum(-1)
            feats = feats - cls.unsqueeze(1) * (s.unsqueeze(-1) * self.global_debias_factor)
        if taps is not None:
            taps['patch_feats'] = feats.clone()
        if self.upsampler is not None:                                               # :368-372
            name, up = self.upsampler
            D = feats.shape[-1]
            f = feats.permute(0, 2, 1).reshape(B, D, fw, fh)
            f = jbu_upsample(up, name, f, img, taps)
            feats = f.reshape(B, D, -1).permute(0, 2, 1)
        feats = feats / feats.norm(dim=-1, keepdim=True)                             # :374
        logits = feats @ self.query_features.T                                       # :375
        if self.cls_token_lambda != 0:                                               # :378-379
            logits = logits + cls_logits.unsqueeze(1) * self.cls_token_lambda
        if self.upsampler is not None:
            w, h = img.shape[-2], img.shape[-1]
        else:
            w, h = fw, fh
        logits = logits.permute(0, 2, 1).reshape(B, -1, w, h)
        size = img.shape[-2:] if logit_size is None else logit_size
        return F.interpolate(logits, size=size, mode='bilinear')                     # :388-391

    def crops(self, img: Tensor):
        """Crop list of forward_slide incl. the pad-to-patch-multiple of :427-431."""
        _, _, H, W = img.shape
        out = []
        for (y1, y2, x1, x2) in slide_windows(H, W, self.slide_stride, self.slide_crop):
            c = img[:, :, y1:y2, x1:x2]
            pad = compute_padsize(c.shape[2], c.shape[3], self.patch)
            if any(pad):
                c = F.pad(c, pad)
            out.append(((y1, y2, x1, x2), pad, c))
        return out

    # segmentor.py:394-451
    def forward_slide(self, img: Tensor, ori_shape=None, crop_logits_out: Optional[list] = None) -> Tensor:
        B, _, H, W = img.shape
        preds = img.new_zeros((B, self.num_queries, H, W))
        count = img.new_zeros((B, 1, H, W))
        for (y1, y2, x1, x2), pad, c in self.crops(img):
            lg = self.forward_feature(c)
            if any(pad):
                l, t = pad[0], pad[2]
                lg = lg[:, :, t:t + (y2 - y1), l:l + (x2 - x1)]
            if crop_logits_out is not None:
                crop_logits_out.append(lg.clone())
            preds[:, :, y1:y2, x1:x2] += lg
            count[:, :, y1:y2, x1:x2] += 1
        assert (count == 0).sum() == 0
        preds = preds / count
        size = (H, W) if ori_shape is None else tuple(ori_shape[:2])
        return F.interpolate(preds, size=size, mode='bilinear')

    # segmentor.py:475-499
    def postprocess(self, seg_logits: Tensor):
        """seg_logits [Q,H,W] (one image).  Returns (probs [K,H,W], pred [1,H,W] int64)."""
        pr = (seg_logits * self.logit_scale).softmax(0)
        if self.num_classes != self.num_queries:
            onehot = F.one_hot(self.query_idx).T.view(self.num_classes, self.num_queries, 1, 1)
            pr = (pr.unsqueeze(0) * onehot).max(1)[0]
        pred = pr.argmax(0, keepdim=True)
        pred[pr.max(0, keepdim=True)[0] < self.prob_thd] = self.bg_idx
        return pr, pred

    # segmentor.py:453-473
    def predict(self, inputs: Tensor, ori_shape=None):
        if self.slide_crop > 0:
            lg = self.forward_slide(inputs, ori_shape)
        else:
            lg = self.forward_feature(inputs, ori_shape if ori_shape is not None else inputs.shape[-2:])
        return self.postprocess(lg[0])


def postprocess_from_crop_logits(crop_logits: Tensor, windows: Sequence[Tuple[int, int, int, int]],
                                 H: int, W: int, query_idx: Sequence[int], logit_scale: float,
                                 prob_thd: float, bg_idx: int, out_size=None):
    """forward_slide accumulation (:440-449) + postprocess_result (:478-489) given per-crop
    logits [n_crops, Q, hc, wc] (already cut to the window size).  Used to check the fused
    accumulate->argmax kernel bit-exactly on labels."""
    n, Q = crop_logits.shape[:2]
    preds = crop_logits.new_zeros((1, Q, H, W))
    count = crop_logits.new_zeros((1, 1, H, W))
    for i, (y1, y2, x1, x2) in enumerate(windows):
        preds[:, :, y1:y2, x1:x2] += crop_logits[i]
        count[:, :, y1:y2, x1:x2] += 1
    preds = preds / count
    if out_size is not None and tuple(out_size) != (H, W):
        preds = F.interpolate(preds, size=tuple(out_size), mode='bilinear')
    qi = torch.tensor(list(query_idx), dtype=torch.int64)
    K = int(qi.max()) + 1
    pr = (preds[0] * logit_scale).softmax(0)
    if K != Q:
        onehot = F.one_hot(qi).T.view(K, Q, 1, 1)
        pr = (pr.unsqueeze(0) * onehot).max(1)[0]
    pred = pr.argmax(0, keepdim=True)
    pred[pr.max(0, keepdim=True)[0] < prob_thd] = bg_idx
    return preds[0], pr, pred


# ----------------------------------------------------------------------------------------------
# mmseg IoUMetric (external dependency mmsegmentation==1.2.2, not vendored in the reference;
# restated from its published algorithm: intersect_and_union + compute_metrics).
# ----------------------------------------------------------------------------------------------

def intersect_and_union(pred: Tensor, label: Tensor, num_classes: int, ignore_index: int = 255):
    """Per-class intersect / pred / label pixel counts (int64) for one image."""
    mask = label != ignore_index
    pred, label = pred[mask], label[mask]
    inter = pred[pred == label]
    ai = torch.bincount(inter, minlength=num_classes)[:num_classes]
    ap = torch.bincount(pred[(pred >= 0) & (pred < num_classes)], minlength=num_classes)[:num_classes]
    al = torch.bincount(label[(label >= 0) & (label < num_classes)], minlength=num_classes)[:num_classes]
    return ai, ap, al


def iou_metrics(ai: Tensor, ap: Tensor, al: Tensor):
    """aAcc / mIoU / mAcc (percent, nan-mean over classes) from summed histograms."""
    ai, ap, al = ai.double(), ap.double(), al.double()
    union = ap + al - ai
    iou = ai / union
    acc = ai / al
    return dict(aAcc=float(ai.sum() / al.sum() * 100),
                mIoU=float(torch.nanmean(iou) * 100),
                mAcc=float(torch.nanmean(acc) * 100))
