"""Container-only harness that imports the UNMODIFIED reference Python from /root/reference.

TEST INFRASTRUCTURE.  Used by ``oracle/gen_golden.py`` (to pin the oracle and to write the
golden vectors under ``tests/golden/``) and by the ``--impl reference`` / ``cpu_baseline`` legs
when the reference tree is present.  /root/reference does not exist on the GPU box, so nothing
in the gpu tests, smoke() or the default bench path imports this module.

Shims: an identity ``ftfy.fix_text``; stub ``mmseg`` / ``mmengine`` / ``BLIP`` /
``gem`` modules so that ``segmentor.py`` imports; ``upsamplers.AdaptiveConv`` replaced by the
tap-loop form of the in-tree ``adaptive_conv_py_simple`` (simfeatup_dev/upsamplers.py:14-25).
"""
import os
import sys
import types

import torch
import torch.nn as nn

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('CLIPSEG_REF', '/root/reference')
if not os.path.isdir(os.path.join(REF, 'open_clip')) and os.path.isdir(os.path.join(_ROOT, 'baseline', '_ref', 'open_clip')):
    REF = os.path.join(_ROOT, 'baseline', '_ref')          # staged by oracle/stage_ref.py for GPU-box runs


def available() -> bool:
    return os.path.isdir(os.path.join(REF, 'open_clip'))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


_installed = False
_CUDA_AUTOCAST = torch.cuda.amp.autocast          # build_ref_segmentor (CPU) substitutes it; the CUDA builder restores it


def install():
    """Install the shims and put the reference on sys.path (idempotent)."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f'reference tree not found at {REF}')
    _stub('ftfy', fix_text=lambda s: s)

    class BaseSegmentor(nn.Module):
        def __init__(self, data_preprocessor=None, init_cfg=None):
            super().__init__()
            self.data_preprocessor = data_preprocessor

    class SegDataPreProcessor(nn.Module):
        def __init__(self, mean=None, std=None, bgr_to_rgb=False, **kw):
            super().__init__()
            self.mean, self.std, self.bgr_to_rgb = mean, std, bgr_to_rgb

    class _Registry:
        def register_module(self, *a, **k):
            return lambda cls: cls

    class PixelData:
        def __init__(self, data=None, **kw):
            self.data = data

    _stub('mmseg'); _stub('mmseg.models')
    _stub('mmseg.models.segmentors', BaseSegmentor=BaseSegmentor)
    _stub('mmseg.models.data_preprocessor', SegDataPreProcessor=SegDataPreProcessor)
    _stub('mmseg.registry', MODELS=_Registry())
    _stub('mmengine'); _stub('mmengine.structures', PixelData=PixelData)
    _stub('BLIP'); _stub('BLIP.models'); _stub('BLIP.models.blip_retrieval', blip_retrieval=None)
    _stub('gem')
    sys.path.insert(0, REF)
    import simfeatup_dev.upsamplers as ups   # noqa

    class _AC:
        @staticmethod
        def apply(inp, filt):
            b, c, h1, w1 = inp.shape
            _, h2, w2, f1, f2 = filt.shape
            # reduced-precision runs (the bf16 yardstick) accumulate the taps in fp32, as a device kernel would
            low = inp.dtype in (torch.float16, torch.bfloat16)
            out = torch.zeros(b, c, h2, w2, dtype=torch.float32 if low else inp.dtype, device=inp.device)
            for i in range(f1):
                for j in range(f2):
                    out.addcmul_(inp[:, :, i:i + h2, j:j + w2].to(out.dtype), filt[:, None, :, :, i, j].to(out.dtype))
            return out.to(inp.dtype)
    ups.AdaptiveConv = _AC
    _installed = True


def build_ref_clip(cfg: dict, state_dict: dict, precision: str = 'fp32'):
    """Reference ``CLIP`` (open_clip/model.py:220) for a config dict, weights loaded strictly."""
    install()
    from open_clip.model import CLIP, convert_weights_to_lp, get_cast_dtype
    v = {k: cfg['vision_cfg'][k] for k in ('image_size', 'layers', 'width', 'patch_size')}
    if 'head_width' in cfg['vision_cfg']:
        v['head_width'] = cfg['vision_cfg']['head_width']
    m = CLIP(embed_dim=cfg['embed_dim'], vision_cfg=v, text_cfg=dict(cfg['text_cfg']),
             quick_gelu=cfg.get('quick_gelu', False), cast_dtype=get_cast_dtype(precision))
    m.load_state_dict(state_dict, strict=True)
    if precision in ('fp16', 'bf16'):
        convert_weights_to_lp(m, torch.float16 if precision == 'fp16' else torch.bfloat16)
    return m.eval()


def build_ref_segmentor_cuda(cfg: dict, state_dict: dict, name_path: str, *, upsampler=None, **kw):
    """The reference's unmodified ``SegmentorEx`` the way it runs on a GPU: fp16 weights (create_model(...,
    precision='fp16')), ``.cuda().half()`` upsampler under cuda autocast (segmentor.py:280,370-371).  Only
    ``create_model`` (no network) and the un-vendored featup ``AdaptiveConv`` op are substituted."""
    install()
    import tempfile
    import segmentor as refseg
    torch.cuda.amp.autocast = _CUDA_AUTOCAST
    refseg.create_model = lambda *a, **k: build_ref_clip(cfg, state_dict, 'fp16')
    extra = {}
    if upsampler is not None:
        name, up_sd = upsampler
        f = tempfile.NamedTemporaryFile(suffix='.ckpt', delete=False)
        torch.save({'state_dict': {'upsampler.' + k: v for k, v in up_sd.items()}}, f.name)   # k[10:] at segmentor.py:282
        extra = dict(apply_sim_feat_up=True, sim_feat_up_cfg=dict(model_name=name, model_path=f.name))
        _orig_load = torch.load
        torch.load = lambda p, *a, **k: _orig_load(p, *a, **{**k, 'weights_only': False})
    cwd = os.getcwd()
    try:
        seg = refseg.SegmentorEx(clip_type='CLIP', vit_type='ViT-B/16', name_path=name_path,
                                 device=torch.device('cuda'), **extra, **kw)
    finally:
        os.chdir(cwd)
        if upsampler is not None:
            torch.load = _orig_load
            os.unlink(f.name)
    return seg


def build_ref_segmentor(cfg: dict, state_dict: dict, name_path: str, *, precision='fp32',
                        upsampler=None, **kw):
    """The reference's unmodified ``SegmentorEx`` (segmentor.py:26) on the CPU.  ``create_model``
    is replaced (no network) by the synthetic-weight CLIP; the upsampler is attached by hand
    because segmentor.py:280 hard-codes ``.cuda().half()``."""
    install()
    import segmentor as refseg
    refseg.create_model = lambda *a, **k: build_ref_clip(cfg, state_dict, precision)
    cwd = os.getcwd()
    try:
        seg = refseg.SegmentorEx(clip_type='CLIP', vit_type='ViT-B/16', name_path=name_path,
                                 device=torch.device('cpu'), apply_sim_feat_up=False, **kw)
    finally:
        os.chdir(cwd)
    if upsampler is not None:
        name, up_sd = upsampler
        from simfeatup_dev.upsamplers import get_upsampler
        seg.feat_dim = seg.query_features.shape[-1]
        seg.upsampler = get_upsampler(name, seg.feat_dim).eval()
        seg.upsampler.load_state_dict(up_sd, strict=True)
        seg.apply_sim_feat_up = True
        # segmentor.py:370-371 wraps the call in torch.cuda.amp.autocast() and .half(); on the CPU
        # fp32 harness both must be identities.
        # (the bf16 yardstick run keeps autocast on, in bf16: CPU autocast has no fp16 kernels for these ops)
        low = precision in ('fp16', 'bf16')
        torch.cuda.amp.autocast = lambda *a, **k: torch.autocast('cpu', dtype=torch.bfloat16, enabled=low)
        _orig = seg.upsampler.forward
        class _NoHalf(torch.Tensor):
            pass
        def fwd(src, g):
            out = _orig(src, g)
            out.half = (lambda: out.to(torch.bfloat16)) if low else (lambda: out)
            return out
        seg.upsampler.forward = fwd
    return seg
