"""``create_model`` with the reference signature (open_clip/factory.py:165-182).

``pretrained=None`` gives deterministic synthetic weights (no network here); a path loads an
open_clip / OpenAI-format state dict; a hub tag ('openai', 'laion2b_s34b_b88k', ...) is resolved in
``$CLIPSEG_WEIGHTS_DIR/<model>-<tag>.pt`` and, failing that, raises like the reference does when the
download is impossible -- unless ``CLIPSEG_SYNTHETIC_WEIGHTS=1`` asks for synthetic weights.
"""
import logging
import os
from typing import Optional, Union

import torch

from .model import CLIP
from .model_configs import get_model_config, list_models
from .synthetic import synthetic_clip_state_dict


def _load_state(path):
    ck = torch.load(path, map_location='cpu', weights_only=False)
    if isinstance(ck, dict) and 'state_dict' in ck:
        ck = ck['state_dict']
    if hasattr(ck, 'state_dict'):
        ck = ck.state_dict()
    sd = {(k[7:] if k.startswith('module.') else k): v for k, v in ck.items()}
    for k in ('input_resolution', 'context_length', 'vocab_size'):
        sd.pop(k, None)
    return sd


def create_model(model_name: str, pretrained: Optional[str] = None, precision: str = 'fp32',
                 device: Union[str, torch.device] = 'cpu', jit: bool = False, force_quick_gelu: bool = False,
                 force_custom_text: bool = False, force_patch_dropout=None, force_image_size=None,
                 force_preprocess_cfg=None, pretrained_image: bool = False, pretrained_hf: bool = True,
                 cache_dir: Optional[str] = None, output_dict=None, require_pretrained: bool = False,
                 synthetic_seed: int = 0, **model_kwargs):
    model_name = model_name.replace('/', '-')                       # factory.py:194
    cfg = get_model_config(model_name)
    if cfg is None:
        logging.error(f'Model config for {model_name} not found; available models {list_models()}.')
        raise RuntimeError(f'Model config for {model_name} not found.')
    if force_quick_gelu or (pretrained and pretrained.lower() == 'openai'):
        cfg['quick_gelu'] = True                                    # model.py:472: the openai route is QuickGELU
    if force_image_size is not None:
        cfg['vision_cfg']['image_size'] = force_image_size
    # reference precisions: fp32 | fp16 | bf16 (+ pure_*).  The CUDA path computes in bf16 (tcgen05) or fp32.
    prec = 'fp32' if precision in ('fp32', 'amp') else 'bf16'
    model = CLIP(cfg['embed_dim'], cfg['vision_cfg'], cfg['text_cfg'], cfg['quick_gelu'], precision=prec)
    sd = None
    if pretrained:
        path = pretrained if os.path.exists(pretrained) else None
        if path is None:
            wdir = os.environ.get('CLIPSEG_WEIGHTS_DIR', cache_dir or '')
            cand = os.path.join(wdir, f'{model_name}-{pretrained}.pt') if wdir else ''
            path = cand if cand and os.path.exists(cand) else None
        if path is not None:
            sd = _load_state(path)
        elif os.environ.get('CLIPSEG_SYNTHETIC_WEIGHTS', '0') == '1':
            logging.warning(f'pretrained={pretrained!r} not available offline: using synthetic weights')
        else:
            raise RuntimeError(f'Pretrained weights ({pretrained}) not found for model {model_name}. '
                               f'Set CLIPSEG_WEIGHTS_DIR, or CLIPSEG_SYNTHETIC_WEIGHTS=1 for synthetic weights.')
    if sd is None:
        if require_pretrained:
            raise RuntimeError(f'Pretrained weights were required for (model: {model_name}, pretrained: '
                               f'{pretrained}) but not loaded.')
        sd = synthetic_clip_state_dict(cfg, synthetic_seed)
    model.load_state_dict(sd, strict=True)
    model.eval()
    return model.to(torch.device(device) if isinstance(device, str) else device)
