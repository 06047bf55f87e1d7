"""``CLIP`` object returned by ``create_model``: a parameter holder whose state-dict keys equal the
reference's (open_clip/model.py:220-254, open_clip/transformer.py:338-444) so real checkpoints load
with ``load_state_dict``; ``encode_image`` and ``encode_text`` (init-time prompt ensemble only,
segmentor.py:157-174) run on the CUDA engines (``VisualEngine`` / ``TextEngine``, engine.py).  A model that lives
on the CPU can still evaluate ``encode_text`` in plain PyTorch: that form exists for the CPU test that pins the
tokenizer + state-dict mapping against the reference golden, never for the segmentation path.
"""
from collections import OrderedDict
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


class _Block(nn.Module):
    def __init__(self, width: int, heads: int, mlp_ratio: float = 4.0):
        super().__init__()
        self.ln_1 = nn.LayerNorm(width)
        self.attn = nn.MultiheadAttention(width, heads)          # holds in_proj_* / out_proj.* only
        self.ln_2 = nn.LayerNorm(width)
        m = int(width * mlp_ratio)
        self.mlp = nn.Sequential(OrderedDict([('c_fc', nn.Linear(width, m)), ('gelu', nn.Identity()),
                                              ('c_proj', nn.Linear(m, width))]))


class _Transformer(nn.Module):
    def __init__(self, width: int, layers: int, heads: int, mlp_ratio: float = 4.0):
        super().__init__()
        self.width, self.layers, self.heads = width, layers, heads
        self.resblocks = nn.ModuleList([_Block(width, heads, mlp_ratio) for _ in range(layers)])


class VisionTower(nn.Module):
    """Holds the ViT parameters; the forward pass is ``VisualEngine`` (engine.py)."""

    def __init__(self, image_size, patch_size, width, layers, heads, mlp_ratio, output_dim, quick_gelu):
        super().__init__()
        self.image_size = (image_size, image_size)
        self.patch_size = (patch_size, patch_size)
        self.grid_size = (image_size // patch_size, image_size // patch_size)
        self.output_dim, self.quick_gelu = output_dim, quick_gelu
        self.conv1 = nn.Conv2d(3, width, patch_size, patch_size, bias=False)
        self.class_embedding = nn.Parameter(torch.zeros(width))
        self.positional_embedding = nn.Parameter(torch.zeros(self.grid_size[0] * self.grid_size[1] + 1, width))
        self.ln_pre = nn.LayerNorm(width)
        self.transformer = _Transformer(width, layers, heads, mlp_ratio)
        self.ln_post = nn.LayerNorm(width)
        self.proj = nn.Parameter(torch.zeros(width, output_dim))
        # attached by the segmentor exactly like the reference does (segmentor.py:216,270)
        self.similarity_enhancer = None
        self.outlier_suppressor = None

    def set_outlier_suppressor(self, suppressor, suppression_layers=None):
        """open_clip/transformer.py:446-469 (only the default layer, layers-2, is supported)."""
        if suppression_layers not in (None, [self.transformer.layers - 2], [-2]):
            raise NotImplementedError('outlier suppression is wired to block layers-2 (transformer.py:609)')
        self.outlier_suppressor = suppressor


class CLIP(nn.Module):
    def __init__(self, embed_dim: int, vision_cfg: dict, text_cfg: dict, quick_gelu: bool = False,
                 precision: str = 'bf16'):
        super().__init__()
        v = vision_cfg
        self.visual = VisionTower(v['image_size'], v['patch_size'], v['width'], v['layers'], v['heads'],
                                  v.get('mlp_ratio', 4.0), embed_dim, quick_gelu)
        t = text_cfg
        self.transformer = _Transformer(t['width'], t['layers'], t['heads'])
        self.context_length, self.vocab_size = t['context_length'], t['vocab_size']
        self.token_embedding = nn.Embedding(t['vocab_size'], t['width'])
        self.positional_embedding = nn.Parameter(torch.zeros(t['context_length'], t['width']))
        self.ln_final = nn.LayerNorm(t['width'])
        self.text_projection = nn.Parameter(torch.zeros(t['width'], embed_dim))
        self.logit_scale = nn.Parameter(torch.zeros([]))
        mask = torch.full((t['context_length'], t['context_length']), float('-inf')).triu_(1)
        self.register_buffer('attn_mask', mask, persistent=False)     # transformer.py:1047-1053
        self.quick_gelu = quick_gelu
        self.embed_dim = embed_dim
        self.vision_cfg = dict(v)
        self.precision = precision          # 'bf16' (tcgen05 path) or 'fp32' (verification mode)
        self._engine = None
        self._text_engine = None
        for p in self.parameters():
            p.requires_grad_(False)

    # ---- engine management ------------------------------------------------------------------
    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        self._engine = None
        self._text_engine = None
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def text_engine(self, device=None):
        """The CUDA text tower over a snapshot of the current text-side weights (built lazily)."""
        from ..engine import TextEngine
        device = torch.device(device if device is not None else self.text_projection.device)
        if device.type != 'cuda':
            raise RuntimeError('clip_decontamination_b200 runs on CUDA only (there is no CPU fallback)')
        if getattr(self, '_text_engine', None) is None or self._text_engine.device != device:
            sd = {k: p.detach() for k, p in self.state_dict().items() if not k.startswith('visual.')}
            self._text_engine = TextEngine(sd, width=self.transformer.width, layers=self.transformer.layers,
                                           heads=self.transformer.heads, quick_gelu=self.quick_gelu,
                                           precision=self.precision, device=device)
        return self._text_engine

    def visual_engine(self, device=None):
        """The CUDA engine over a snapshot of the current ``visual.*`` weights (built lazily)."""
        from ..engine import VisualEngine
        if device is None:
            device = self.visual.proj.device
        device = torch.device(device)
        if device.type != 'cuda':
            raise RuntimeError('clip_decontamination_b200 runs on CUDA only: move the model to a B200 '
                               '(there is no CPU fallback)')
        if self._engine is None or self._engine.device != device:
            v = self.vision_cfg
            sd = {k: p.detach() for k, p in self.visual.state_dict().items()}
            self._engine = VisualEngine(sd, width=v['width'], layers=v['layers'], heads=v['heads'],
                                        patch_size=v['patch_size'], image_size=v['image_size'],
                                        embed_dim=self.embed_dim, quick_gelu=self.quick_gelu,
                                        precision=self.precision, device=device)
        return self._engine

    # ---- open_clip/model.py:265-286 ---------------------------------------------------------
    @torch.no_grad()
    def encode_image(self, image, model_type, ignore_residual: bool = False, output_cls_token: bool = False,
                     normalize: bool = False, apply_layer_fusion: bool = False, layer_fusion_lambda: float = 0.5,
                     layer_fusion_threshold: float = 0.7, apply_similarity_enhancement: bool = False):
        if apply_layer_fusion:
            raise NotImplementedError('apply_layer_fusion is outside the hot path (off in every config)')
        eng = self.visual_engine(image.device)
        B, _, H, W = image.shape
        ps = self.visual.patch_size[0]
        ch, cw = (H // ps) * ps, (W // ps) * ps
        tall = image.detach().float().permute(1, 0, 2, 3).reshape(3, B * H, W).contiguous()
        wins = torch.tensor([(b * H, 0, ch, cw) for b in range(B)], dtype=torch.int32, device=image.device)
        se, osup = self.visual.similarity_enhancer, self.visual.outlier_suppressor
        sim_cfg = None
        if se is not None and apply_similarity_enhancement:
            sim_cfg = dict(similarity_weight=se.similarity_weight, temperature=se.temperature,
                           add_self_similarity=se.add_self_similarity)
        out_cfg = None
        if osup is not None:
            out_cfg = dict(top_k=osup.top_k, contamination_temp=osup.contamination_temp)
        tok, L = eng.encode(tall, wins, ch, cw, 0, 0, model_type, ignore_residual, sim_cfg, out_cfg)
        tok = tok.view(B, L, -1)
        cls, feats = tok[:, 0].clone(), tok[:, 1:].clone()
        if normalize:
            cls, feats = F.normalize(cls, dim=-1), F.normalize(feats, dim=-1)
        return (cls, feats) if output_cls_token else feats

    # ---- open_clip/model.py:288-306 (init-time only) ------------------------------------------
    @torch.no_grad()
    def encode_text(self, text, normalize: bool = False):
        if text.is_cuda:                                              # libclipseg kernels (TextEngine)
            x = self.text_engine(text.device).encode(text)
            return F.normalize(x, dim=-1) if normalize else x
        return self.encode_text_torch(text, normalize)

    @torch.no_grad()
    def encode_text_torch(self, text, normalize: bool = False):
        """Plain-PyTorch text tower for a model held on the CPU (CPU checker of the tokenizer / weight mapping)."""
        x = self.token_embedding(text).float() + self.positional_embedding.float()
        heads = self.transformer.heads
        mask = self.attn_mask.to(x.device)
        for blk in self.transformer.resblocks:
            h = blk.ln_1(x)
            B, Lt, d = h.shape
            qkv = F.linear(h, blk.attn.in_proj_weight, blk.attn.in_proj_bias)
            q, k, v = [t.view(B, Lt, heads, d // heads).permute(0, 2, 1, 3) for t in qkv.chunk(3, dim=-1)]
            a = (q * (d // heads) ** -0.5) @ k.transpose(-1, -2) + mask
            o = (a.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(B, Lt, d)
            x = x + blk.attn.out_proj(o)
            h = blk.mlp.c_fc(blk.ln_2(x))
            h = h * torch.sigmoid(1.702 * h) if self.quick_gelu else F.gelu(h)
            x = x + blk.mlp.c_proj(h)
        x = self.ln_final(x)
        x = x[torch.arange(x.shape[0]), text.argmax(dim=-1)] @ self.text_projection   # EOT token
        return F.normalize(x, dim=-1) if normalize else x
