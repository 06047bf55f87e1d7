"""Deterministic synthetic ("random-init") weights with the reference's state-dict key names.

There is no network for checkpoints, so benchmarks and parity tests run on synthetic weights.
Each tensor is drawn from its own generator seeded by (seed, key name), which makes the values
independent of construction order and identical in this package, in the oracle and in the
reference model (they are loaded into the reference's ``CLIP`` with ``load_state_dict`` when
the golden vectors are generated, see ``oracle/gen_golden.py``).  Scales follow what
``open_clip.create_model(pretrained=None)`` produces (open_clip/transformer.py:372-379,442 and
torch's Linear / MultiheadAttention defaults), except that LayerNorm affine parameters are
perturbed so that a kernel that drops gamma/beta cannot pass a parity test.
"""
import hashlib

import torch


def _gen(seed: int, key: str) -> torch.Generator:
    h = hashlib.sha256(f'{seed}:{key}'.encode()).digest()
    g = torch.Generator(device='cpu')
    g.manual_seed(int.from_bytes(h[:8], 'little') % (2 ** 63))
    return g


def _randn(seed, key, shape, std=1.0, mean=0.0):
    return torch.randn(shape, generator=_gen(seed, key), dtype=torch.float32) * std + mean


def _tower(sd, seed, prefix, width, layers, mlp_ratio=4.0):
    mlp = int(width * mlp_ratio)
    for i in range(layers):
        p = f'{prefix}transformer.resblocks.{i}.'
        sd[p + 'ln_1.weight'] = _randn(seed, p + 'ln_1.weight', (width,), 0.05, 1.0)
        sd[p + 'ln_1.bias'] = _randn(seed, p + 'ln_1.bias', (width,), 0.02)
        sd[p + 'attn.in_proj_weight'] = _randn(seed, p + 'attn.in_proj_weight', (3 * width, width),
                                               (2.0 / (4 * width)) ** 0.5 * 1.2)
        sd[p + 'attn.in_proj_bias'] = _randn(seed, p + 'attn.in_proj_bias', (3 * width,), 0.02)
        sd[p + 'attn.out_proj.weight'] = _randn(seed, p + 'attn.out_proj.weight', (width, width),
                                                width ** -0.5 * 0.6)
        sd[p + 'attn.out_proj.bias'] = _randn(seed, p + 'attn.out_proj.bias', (width,), 0.02)
        sd[p + 'ln_2.weight'] = _randn(seed, p + 'ln_2.weight', (width,), 0.05, 1.0)
        sd[p + 'ln_2.bias'] = _randn(seed, p + 'ln_2.bias', (width,), 0.02)
        sd[p + 'mlp.c_fc.weight'] = _randn(seed, p + 'mlp.c_fc.weight', (mlp, width), width ** -0.5 * 0.6)
        sd[p + 'mlp.c_fc.bias'] = _randn(seed, p + 'mlp.c_fc.bias', (mlp,), 0.02)
        sd[p + 'mlp.c_proj.weight'] = _randn(seed, p + 'mlp.c_proj.weight', (width, mlp), mlp ** -0.5 * 0.6)
        sd[p + 'mlp.c_proj.bias'] = _randn(seed, p + 'mlp.c_proj.bias', (width,), 0.01)


def synthetic_clip_state_dict(cfg: dict, seed: int = 0, text_tower: bool = True) -> dict:
    """State dict for the reference's ``CLIP`` module (open_clip/model.py:220-254) built from a
    config of ``model_configs.get_model_config``."""
    v, t, D = cfg['vision_cfg'], cfg['text_cfg'], cfg['embed_dim']
    w, ps = v['width'], v['patch_size']
    g = v['image_size'] // ps
    sd = {}
    sd['visual.class_embedding'] = _randn(seed, 'visual.class_embedding', (w,), w ** -0.5)
    sd['visual.positional_embedding'] = _randn(seed, 'visual.positional_embedding', (g * g + 1, w), w ** -0.5)
    sd['visual.proj'] = _randn(seed, 'visual.proj', (w, D), w ** -0.5)
    sd['visual.conv1.weight'] = _randn(seed, 'visual.conv1.weight', (w, 3, ps, ps), (3 * ps * ps) ** -0.5 * 0.6)
    for n in ('ln_pre', 'ln_post'):
        sd[f'visual.{n}.weight'] = _randn(seed, f'visual.{n}.weight', (w,), 0.05, 1.0)
        sd[f'visual.{n}.bias'] = _randn(seed, f'visual.{n}.bias', (w,), 0.02)
    _tower(sd, seed, 'visual.', w, v['layers'], v.get('mlp_ratio', 4.0))
    if text_tower:
        tw = t['width']
        sd['positional_embedding'] = _randn(seed, 'positional_embedding', (t['context_length'], tw), 0.01)
        sd['text_projection'] = _randn(seed, 'text_projection', (tw, D), tw ** -0.5)
        sd['logit_scale'] = torch.tensor(2.6592600345611572)
        sd['token_embedding.weight'] = _randn(seed, 'token_embedding.weight', (t['vocab_size'], tw), 0.02)
        sd['ln_final.weight'] = _randn(seed, 'ln_final.weight', (tw,), 0.05, 1.0)
        sd['ln_final.bias'] = _randn(seed, 'ln_final.bias', (tw,), 0.02)
        _tower(sd, seed, '', tw, t['layers'])
    return sd


def synthetic_jbu_state_dict(name: str, feat_dim: int, seed: int = 1, key_dim: int = 32) -> dict:
    """State dict for JBUOne ('up.*') / JBUStack ('up1..4.*') + 'fixup_proj.1.*'
    (simfeatup_dev/upsamplers.py:202-228,278-312).  range_temp / sigma_spatial are drawn from the
    range of the shipped jbu_stack checkpoints (0.05-0.60 / 0.66-1.29)."""
    if name == 'jbu_one':
        mods = [('up.', 5)]
    elif name == 'jbu_stack':
        mods = [(f'up{i}.', 3) for i in range(1, 5)]
    else:
        raise ValueError(f"Unknown upsampler {name}")
    sd = {}
    for i, (p, r) in enumerate(mods):
        d2 = (2 * r + 1) ** 2
        sd[p + 'range_temp'] = torch.tensor(0.25 + 0.08 * i)
        sd[p + 'sigma_spatial'] = torch.tensor(0.9 + 0.07 * i)
        sd[p + 'range_proj.0.weight'] = _randn(seed, p + 'range_proj.0.weight', (key_dim, 3, 1, 1), 0.5)
        sd[p + 'range_proj.0.bias'] = _randn(seed, p + 'range_proj.0.bias', (key_dim,), 0.1)
        sd[p + 'range_proj.3.weight'] = _randn(seed, p + 'range_proj.3.weight', (key_dim, key_dim, 1, 1), key_dim ** -0.5)
        sd[p + 'range_proj.3.bias'] = _randn(seed, p + 'range_proj.3.bias', (key_dim,), 0.1)
        sd[p + 'fixup_proj.0.weight'] = _randn(seed, p + 'fixup_proj.0.weight', (d2, d2 + 3, 1, 1), (d2 + 3) ** -0.5)
        sd[p + 'fixup_proj.0.bias'] = _randn(seed, p + 'fixup_proj.0.bias', (d2,), 0.05)
        sd[p + 'fixup_proj.3.weight'] = _randn(seed, p + 'fixup_proj.3.weight', (d2, d2, 1, 1), d2 ** -0.5 * 0.3)
        sd[p + 'fixup_proj.3.bias'] = _randn(seed, p + 'fixup_proj.3.bias', (d2,), 0.01)
    sd['fixup_proj.1.weight'] = _randn(seed, 'fixup_proj.1.weight', (feat_dim, feat_dim, 1, 1), feat_dim ** -0.5)
    sd['fixup_proj.1.bias'] = _randn(seed, 'fixup_proj.1.bias', (feat_dim,), 0.02)
    return sd
