"""Model configurations for the CLIP families the segmentor can name.

Values follow the json files of the reference's vendored open_clip
(open_clip/model_configs/ViT-B-16.json, ViT-B-16-quickgelu.json, ViT-B-32.json, ViT-L-14.json,
ViT-L-14-quickgelu.json, ViT-H-14.json); `quick_gelu` selects x*sigmoid(1.702x) instead of
erf-GELU (open_clip/model.py:116).
"""
import copy

_V = dict
MODEL_CONFIGS = {
    'ViT-B-16': dict(embed_dim=512,
                     vision_cfg=_V(image_size=224, layers=12, width=768, patch_size=16),
                     text_cfg=_V(context_length=77, vocab_size=49408, width=512, heads=8, layers=12)),
    'ViT-B-32': dict(embed_dim=512,
                     vision_cfg=_V(image_size=224, layers=12, width=768, patch_size=32),
                     text_cfg=_V(context_length=77, vocab_size=49408, width=512, heads=8, layers=12)),
    'ViT-L-14': dict(embed_dim=768,
                     vision_cfg=_V(image_size=224, layers=24, width=1024, patch_size=14),
                     text_cfg=_V(context_length=77, vocab_size=49408, width=768, heads=12, layers=12)),
    'ViT-H-14': dict(embed_dim=1024,
                     vision_cfg=_V(image_size=224, layers=32, width=1280, head_width=80, patch_size=14),
                     text_cfg=_V(context_length=77, vocab_size=49408, width=1024, heads=16, layers=24)),
    # tiny configuration used by unit tests and smoke runs (not a reference config)
    'ViT-tiny-16': dict(embed_dim=64,
                        vision_cfg=_V(image_size=224, layers=4, width=128, patch_size=16),
                        text_cfg=_V(context_length=77, vocab_size=49408, width=64, heads=2, layers=2)),
}
for _n in ('ViT-B-16', 'ViT-B-32', 'ViT-L-14', 'ViT-H-14'):
    _c = copy.deepcopy(MODEL_CONFIGS[_n])
    _c['quick_gelu'] = True
    MODEL_CONFIGS[_n + '-quickgelu'] = _c


def get_model_config(name: str):
    """open_clip/factory.py:194 normalises '/' to '-' before the lookup."""
    name = name.replace('/', '-')
    if name not in MODEL_CONFIGS:
        return None
    cfg = copy.deepcopy(MODEL_CONFIGS[name])
    v = cfg['vision_cfg']
    v.setdefault('head_width', 64)
    v.setdefault('mlp_ratio', 4.0)
    v['heads'] = v['width'] // v['head_width']
    cfg.setdefault('quick_gelu', False)
    return cfg


def list_models():
    return sorted(MODEL_CONFIGS)
