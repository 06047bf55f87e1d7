"""CPU suite (`-m "not gpu"`): the oracle against the golden vectors produced by the unmodified reference,
the host-side logic, and the C ABI surface (the library must load without a GPU and export every symbol
declared in include/clipseg.h; no compute calls here)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import clipseg_oracle as O
from clip_decontamination_b200 import synth
from clip_decontamination_b200.open_clip.model_configs import get_model_config
from clip_decontamination_b200.open_clip.synthetic import synthetic_clip_state_dict, synthetic_jbu_state_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 2e-5   # oracle vs reference (both fp32 CPU; differences are summation order only)


def _vis(cfg):
    sd = synthetic_clip_state_dict(cfg, 0, text_tower=False)
    return {k[len('visual.'):]: v for k, v in sd.items() if k.startswith('visual.')}


def _two_crops():
    img = torch.from_numpy(synth.preprocess(synth.voronoi_scene(224, 448, 5)))
    return torch.stack([img[:, :, :224], img[:, :, 224:]])


# ---------------------------------------------------------------- oracle vs golden (reference outputs) --
def test_oracle_vit_tiny_all_variants(gold):
    g = gold('vit_tiny')
    cfg = get_model_config('ViT-tiny-16')
    v, vis, x = cfg['vision_cfg'], _vis(cfg), _two_crops()
    kw = dict(layers=v['layers'], heads=v['heads'], patch=v['patch_size'])
    with torch.no_grad():
        for mt in ['Experimental', 'SCLIP', 'ClearCLIP', 'SFP', 'vanilla', 'SegEarth', 'MaskCLIP']:
            c, t = O.vit_dense_forward(vis, x, model_type=mt, sim_cfg={}, outlier_cfg={'top_k': 30}, **kw)
            assert np.abs(t.numpy() - g[f'{mt}_tokens']).max() < TOL
            assert np.abs(c.numpy() - g[f'{mt}_cls']).max() < TOL
        c, t = O.vit_dense_forward(vis, x, **kw)
        assert np.abs(t.numpy() - g['plain_tokens']).max() < TOL
        c, t = O.vit_dense_forward(vis, x, ignore_residual=False, **kw)
        assert np.abs(t.numpy() - g['residual_tokens']).max() < TOL
        c, t = O.vit_dense_forward(vis, x, quick_gelu=True, **kw)
        assert np.abs(t.numpy() - g['quickgelu_tokens']).max() < TOL


def test_oracle_vit_b16_crop(gold):
    g = gold('vit_b16_crop')
    cfg = get_model_config('ViT-B-16')
    v = cfg['vision_cfg']
    with torch.no_grad():
        c, t = O.vit_dense_forward(_vis(cfg), _two_crops()[:1], layers=12, heads=12, patch=16, sim_cfg={},
                                   outlier_cfg={'top_k': 30})
    assert np.abs(t.numpy() - g['Experimental_tokens']).max() < TOL


@pytest.mark.parametrize('name', ['jbu_one', 'jbu_stack'])
def test_oracle_jbu(gold, name):
    g = gold(f'{name}_c32')
    img = torch.from_numpy(synth.preprocess(synth.voronoi_scene(224, 224, 5)))[None]
    taps = {}
    with torch.no_grad():
        out = O.jbu_upsample(synthetic_jbu_state_dict(name, 32, 1), name, torch.from_numpy(g['source']), img, taps)
    assert np.abs(out.numpy()[:, :, ::4, ::4] - g['out']).max() < TOL
    assert np.abs(taps['jbu_stages'][0].numpy() - g['stage0']).max() < TOL
    assert abs(float(out.double().sum()) - float(g['out_sum'])) < 1e-2


def test_oracle_jbu_stack_real_checkpoint(gold):
    g = gold('jbu_stack_real')
    sd = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('w.')}
    img = torch.from_numpy(synth.preprocess(synth.voronoi_scene(224, 224, 5)))[None]
    src = torch.randn(1, 512, 14, 14, generator=torch.Generator().manual_seed(7))
    with torch.no_grad():
        out = O.jbu_upsample(sd, 'jbu_stack', src, img)
    assert np.abs(out.numpy()[:, ::16, ::4, ::4] - g['out']).max() < TOL


def test_oracle_postprocess_bit_exact(gold):
    g = gold('postproc')
    for tag in ('potsdam', 'loveda', 'road'):
        H, W, thd, bg, stride, crop = g[f'{tag}_meta']
        cl = torch.from_numpy(g[f'{tag}_crop_logits']).float()
        wins = O.slide_windows(int(H), int(W), int(stride), int(crop))
        _, _, pred = O.postprocess_from_crop_logits(cl, wins, int(H), int(W), g[f'{tag}_query_idx'].tolist(), 50,
                                                    float(thd), int(bg))
        assert np.array_equal(pred[0].numpy().astype(np.uint8), g[f'{tag}_labels'])


@pytest.mark.parametrize('name,model,ups', [('seg_tiny_jbu', 'ViT-tiny-16', 'jbu_one'),
                                            ('seg_potsdam_noup', 'ViT-B-16', None)])
def test_oracle_full_segmentor(gold, name, model, ups):
    g = gold(name)
    cfg = get_model_config(model)
    v = cfg['vision_cfg']
    H, W, thd, bg, seed = int(g['meta'][0]), int(g['meta'][1]), float(g['meta'][2]), int(g['meta'][3]), int(g['meta'][4])
    up = (ups, synthetic_jbu_state_dict(ups, cfg['embed_dim'], 1)) if ups else None
    orc = O.SegOracle(_vis(cfg), torch.from_numpy(g['query_features']), g['query_idx'].tolist(), layers=v['layers'],
                      heads=v['heads'], patch=v['patch_size'], prob_thd=thd, bg_idx=bg, global_debias_factor=0.2,
                      upsampler=up, sim_cfg={}, outlier_cfg={'top_k': 30})
    img = torch.from_numpy(synth.preprocess(synth.voronoi_scene(H, W, seed)))[None]
    with torch.no_grad():
        _, pred = orc.predict(img)
        lg = orc.forward_slide(img)
    assert np.abs(lg[0].numpy()[:, ::4, ::4] - g['logits_sub']).max() < 1e-5
    assert (pred[0].numpy() == g['labels']).mean() >= 0.9999


def test_oracle_whole_image_long_sequence(gold):
    """slide_crop = 0 (segmentor.py:470-471) on a 400x400 image: ONE crop of L = 626 tokens, interpolated positional
    embedding, a single bilinear resize to ori_shape.  The oracle against the unmodified reference's golden."""
    g = gold('seg_whole_400_noup')
    cfg = get_model_config('ViT-B-16')
    v = cfg['vision_cfg']
    m = g['meta']
    H, W, thd, bg, seed, crop = int(m[0]), int(m[1]), float(m[2]), int(m[3]), int(m[4]), int(m[6])
    assert crop == 0
    orc = O.SegOracle(_vis(cfg), torch.from_numpy(g['query_features']), g['query_idx'].tolist(), layers=v['layers'],
                      heads=v['heads'], patch=v['patch_size'], prob_thd=thd, bg_idx=bg, slide_crop=0,
                      global_debias_factor=0.2, upsampler=None, sim_cfg={}, outlier_cfg={'top_k': 30})
    img = torch.from_numpy(synth.preprocess(synth.voronoi_scene(H, W, seed)))[None]
    with torch.no_grad():
        lg = orc.forward_feature(img, (H, W))
        _, pred = orc.postprocess(lg[0])
    sub = int(m[10])
    assert np.abs(lg[0].numpy()[:, ::sub, ::sub] - g['logits_sub'].astype(np.float32)).max() < 1e-5
    assert (pred.numpy().reshape(H, W) == g['labels']).mean() >= 0.9999


def test_oracle_iou_metrics():
    pred = torch.tensor([0, 1, 1, 2, 2, 2, 0, 255 % 3])
    lab = torch.tensor([0, 1, 2, 2, 2, 255, 1, 0])
    ai, ap, al = O.intersect_and_union(pred, lab, 3)
    assert ai.tolist() == [2, 1, 2] and ap.tolist() == [3, 2, 2] and al.tolist() == [2, 2, 3]
    m = O.iou_metrics(ai, ap, al)
    assert abs(m['aAcc'] - 5 / 7 * 100) < 1e-9 and abs(m['mIoU'] - (2 / 3 + 1 / 3 + 2 / 3) / 3 * 100) < 1e-9


# ---------------------------------------------------------------- host logic -----------------------
def test_windows_and_padding_match_reference_rules():
    from clip_decontamination_b200.engine import slide_windows, compute_padsize
    assert [w[0] for w in slide_windows(512, 512, 112, 224)][::4] == [0, 112, 224, 288]       # snapped last window
    assert len(slide_windows(1024, 1024, 112, 224)) == 81 and slide_windows(1024, 1024, 112, 224)[-1] == (800, 800, 224, 224)
    assert len(slide_windows(896, 896, 112, 224)) == 49 and len(slide_windows(448, 448, 112, 224)) == 9
    assert slide_windows(200, 300, 112, 224) == [(0, 0, 200, 224), (0, 76, 200, 224)]
    for (H, W) in [(512, 512), (300, 260), (1300, 1301), (100, 90)]:
        ours = [(y, y + h, x, x + w) for (y, x, h, w) in slide_windows(H, W, 112, 224)]
        assert ours == O.slide_windows(H, W, 112, 224)
        cnt = np.zeros((H, W), int)
        for (y1, y2, x1, x2) in ours:
            cnt[y1:y2, x1:x2] += 1
        assert cnt.min() >= 1
    assert compute_padsize(200, 224, 16) == (0, 0, 4, 4) and compute_padsize(199, 154, 16) == (3, 3, 4, 5)
    assert compute_padsize(224, 224, 14) == (0, 0, 0, 0) == O.compute_padsize(224, 224, 14)


def test_get_cls_idx_semantics(tmp_path):
    from clip_decontamination_b200.segmentor import get_cls_idx
    names, idx = get_cls_idx(os.path.join(ROOT, 'configs', 'cls_potsdam.txt'))
    assert names == ['road', 'parking lot', 'building', 'low vegetation', 'tree', 'car', 'clutter', 'background']
    assert idx == [0, 0, 1, 2, 3, 4, 5, 5]
    p = tmp_path / 'c.txt'
    p.write_text('a, b\nc')                     # no whitespace stripping except the newline (segmentor.py:618-621)
    assert get_cls_idx(str(p)) == (['a', ' b', 'c'], [0, 0, 1])
    assert get_cls_idx(str(p)) == O.get_cls_idx(str(p))


def test_shard_indices_match_mmengine_default_sampler():
    from clip_decontamination_b200.dist import shard_indices, owned_mask
    for n, world in [(10, 4), (8, 8), (3, 8), (17, 2), (1, 1)]:
        seen = []
        for r in range(world):
            idx, own = shard_indices(n, r, world), owned_mask(n, r, world)
            assert len(idx) == (n + world - 1) // world == len(own)
            seen += [i for i, o in zip(idx, own) if o]
        assert sorted(seen) == list(range(n))
    assert shard_indices(10, 1, 4) == [1, 5, 9] and shard_indices(10, 3, 4) == [3, 7, 1]     # wrap padding


def test_model_factory_and_state_dict_names():
    from clip_decontamination_b200.open_clip import create_model
    m = create_model('ViT-B/16', pretrained=None)
    keys = set(m.state_dict().keys())
    for k in ('visual.conv1.weight', 'visual.class_embedding', 'visual.positional_embedding', 'visual.proj',
              'visual.transformer.resblocks.11.attn.in_proj_weight', 'visual.transformer.resblocks.0.mlp.c_proj.bias',
              'visual.ln_post.weight', 'token_embedding.weight', 'text_projection', 'logit_scale',
              'transformer.resblocks.0.attn.out_proj.weight', 'ln_final.bias', 'positional_embedding'):
        assert k in keys, k
    assert m.visual.patch_size == (16, 16)
    with pytest.raises(RuntimeError, match='not found'):
        create_model('ViT-Nope-1')
    with pytest.raises(RuntimeError):
        m.encode_image(torch.zeros(1, 3, 224, 224), 'Experimental', True)       # CPU tensor: no fallback


def test_upsampler_factory_loads_reference_key_names(gold):
    from clip_decontamination_b200.simfeatup_dev.upsamplers import get_upsampler
    g = gold('jbu_stack_real')
    sd = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('w.')}
    get_upsampler('jbu_stack', 512).load_state_dict(sd, strict=True)
    get_upsampler('jbu_one', 64).load_state_dict(synthetic_jbu_state_dict('jbu_one', 64, 1), strict=True)
    with pytest.raises(ValueError, match='Unknown upsampler'):
        get_upsampler('nope', 8)


def test_tokenizer_and_text_tower_match_reference(gold):
    from clip_decontamination_b200.open_clip import create_model, tokenizer
    if tokenizer.find_bpe_vocab() is None:
        pytest.skip('CLIP BPE vocabulary not available on this machine')
    g = gold('text_tiny')
    toks = tokenizer.tokenize([str(s) for s in g['prompts']])
    assert np.array_equal(toks.numpy(), g['tokens'])
    f = create_model('ViT-tiny-16', None).encode_text(toks)
    assert np.abs(f.numpy() - g['feats']).max() < 1e-5


def test_prompt_templates(gold):
    from clip_decontamination_b200.prompts.imagenet_template import openai_imagenet_template
    g = gold('text_tiny')
    assert [t('X') for t in openai_imagenet_template] == [str(s) for s in g['template_probe']]


# ---------------------------------------------------------------- C ABI surface ---------------------
def test_cabi_exports_every_declared_symbol():
    from clip_decontamination_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'clipseg.h')).read()
    declared = set(re.findall(r'CSEG_API\s+(?:int|long long)\s+(cseg_\w+)\s*\(', hdr))
    assert len(declared) >= 20
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    out = subprocess.run(['nm', '-D', '--defined-only', _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r' T (cseg_\w+)', out))
    assert declared <= exported, declared - exported
    assert _lib.lib.cseg_version() == 200


def test_cseg_image_struct_layout_matches_ctypes(tmp_path):
    """include/clipseg.h is valid C and struct cseg_image has the layout the ctypes binding assumes."""
    from clip_decontamination_b200 import _lib
    src = tmp_path / 'layout.c'
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "clipseg.h"\nint main(void){'
                   'printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(cseg_image), offsetof(cseg_image, dtype), '
                   'offsetof(cseg_image, img_h), offsetof(cseg_image, stride_img), offsetof(cseg_image, stride_x), '
                   'offsetof(cseg_image, chan), offsetof(cseg_image, mean), offsetof(cseg_image, std)); return 0;}')
    exe = tmp_path / 'layout'
    subprocess.run(['gcc', '-std=c99', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    CI = _lib.CsegImage
    assert got == [C.sizeof(CI), CI.dtype.offset, CI.img_h.offset, CI.stride_img.offset, CI.stride_x.offset,
                   CI.chan.offset, CI.mean.offset, CI.std.offset], got


def test_cabi_argument_validation_without_gpu():
    """argument errors are reported through the return code + cseg_last_error before any CUDA call."""
    from clip_decontamination_b200 import _lib
    rc = _lib.lib.cseg_accum_argmax(None, 0, 4, 1, 1, 1, 1, 0, 0, None, 0, 0, 0, 0, None, 1, 50.0, 0.0, 0, None, None,
                                    None, None)
    assert rc == -1 and 'empty' in _lib.last_error()
    rc = _lib.lib.cseg_attention(0, None, 1, 1, 12, 64, 0, None, 1.0, None, None, None)
    assert rc == -1 and 'L=1' in _lib.last_error()
    # any sequence length is accepted up to the one-score-row-per-warp limit of the long-sequence kernel (51 200 tokens)
    rc = _lib.lib.cseg_attention(0, None, 1, 60000, 12, 64, 0, None, 1.0, None, None, None)
    assert rc == -1 and 'L=60000' in _lib.last_error()
    # the layout-1 similarity map of the header and of the binding agree
    from clip_decontamination_b200 import ops
    hdr = open(os.path.join(ROOT, 'include', 'clipseg.h')).read()
    assert int(re.search(r'#define CSEG_SIMT_COLS (\d+)', hdr).group(1)) == ops.SIMT_COLS
    assert int(re.search(r'#define CSEG_SIMT_COLS_MAX (\d+)', hdr).group(1)) == ops.SIMT_COLS_MAX
    assert ops.simt_floats(197) == 7 * 208 * 32 and ops.simt_floats(257) == 9 * 272 * 32


# ---------------------------------------------------------------- N > 1 path on CPU (gloo, world 2) --
def test_histogram_allreduce_world2_gloo(tmp_path):
    script = tmp_path / 'w.py'
    script.write_text(f'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {ROOT!r})
from clip_decontamination_b200.dist import shard_indices, owned_mask, allreduce_hist, iou_metrics
from oracle import clipseg_oracle as O
from clip_decontamination_b200 import synth
dist.init_process_group('gloo')
r, w = dist.get_rank(), dist.get_world_size()
K, n = 6, 5
hist = torch.zeros(3, K, dtype=torch.int64)
for i, own in zip(shard_indices(n, r, w), owned_mask(n, r, w)):
    if not own: continue
    pred = torch.from_numpy(synth.synthetic_labels(64, 48, K, 10 + i)).long() % K
    lab = torch.from_numpy(synth.synthetic_labels(64, 48, K, 20 + i)).long()
    hist += torch.stack(O.intersect_and_union(pred, lab, K))
allreduce_hist(hist)
ref = torch.zeros(3, K, dtype=torch.int64)
for i in range(n):
    pred = torch.from_numpy(synth.synthetic_labels(64, 48, K, 10 + i)).long() % K
    lab = torch.from_numpy(synth.synthetic_labels(64, 48, K, 20 + i)).long()
    ref += torch.stack(O.intersect_and_union(pred, lab, K))
assert torch.equal(hist, ref), (hist, ref)
m = iou_metrics(hist); o = O.iou_metrics(*ref)
assert all(abs(m[k] - o[k]) < 1e-9 for k in o)
dist.destroy_process_group()
print('ok', r)
''')
    res = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                          '--master-addr', '127.0.0.1', '--master-port', '29631', str(script)],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.count('ok') == 2


def test_bench_workloads_are_runnable_configs():
    """every --workload of bench.py names an existing class file, a known model and query features of that model's width"""
    import importlib
    import numpy as np
    sys.path.insert(0, ROOT)
    bench = importlib.import_module('bench')
    from clip_decontamination_b200.open_clip.model_configs import get_model_config
    assert bench.WORKLOADS['vaihingen512']['H'] == 512 and 'loveda1024_vitl' in bench.WORKLOADS
    for name, wl in bench.WORKLOADS.items():
        cfg = get_model_config(wl['model'].replace('/', '-'))
        assert os.path.exists(os.path.join(ROOT, 'configs', f"cls_{wl['cls']}.txt")), name
        if wl.get('text'):
            qf = np.load(os.path.join(ROOT, 'tests', 'golden', wl['text']))['query_features']
        else:
            qf = np.load(os.path.join(ROOT, 'tests', 'golden', 'bench_text.npz'))[f"{wl['cls']}_query_features"]
        assert qf.shape[1] == cfg['embed_dim'], (name, qf.shape)
        assert wl['up'] is False or cfg['vision_cfg']['patch_size'] == 16, name     # JBU x16 needs patch 16 (segmentor.py:372)
