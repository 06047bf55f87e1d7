"""End-to-end parity of the CUDA engines against the golden vectors produced by the UNMODIFIED reference
(tests/golden/*.npz, see oracle/gen_golden.py) and against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): logits within 1e-4 abs in the fp32 mode and 1e-2 abs in bf16;
label agreement >= 99.9 % in the fp32 mode; in bf16 the agreement is reported next to the top-2 margin
(the reference's own bf16-vs-fp32 agreement on such scenes is ~97 %) and must hold on the
pixels whose oracle margin exceeds twice the measured max logit error."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import clipseg_oracle as O  # noqa: E402
from clip_decontamination_b200 import synth  # noqa: E402
from clip_decontamination_b200.open_clip.model_configs import get_model_config  # noqa: E402
from clip_decontamination_b200.open_clip.synthetic import synthetic_clip_state_dict, synthetic_jbu_state_dict  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _visual_sd(cfg, seed=0):
    sd = synthetic_clip_state_dict(cfg, seed, text_tower=False)
    return {k[len('visual.'):]: v for k, v in sd.items() if k.startswith('visual.')}


def _visual_engine(name, precision):
    from clip_decontamination_b200.engine import VisualEngine
    cfg = get_model_config(name)
    v = cfg['vision_cfg']
    eng = VisualEngine(_visual_sd(cfg), width=v['width'], layers=v['layers'], heads=v['heads'],
                       patch_size=v['patch_size'], image_size=v['image_size'], embed_dim=cfg['embed_dim'],
                       quick_gelu=cfg['quick_gelu'], precision=precision)
    return cfg, eng


def _two_crops_img():
    return torch.from_numpy(synth.preprocess(synth.voronoi_scene(224, 448, 5)))


def _encode(eng, n, **kw):
    img = _two_crops_img().cuda()
    wins = torch.tensor([(0, 0, 224, 224), (0, 224, 224, 224)][:n], dtype=torch.int32).cuda()
    taps = {}
    tok, L = eng.encode(img, wins, 224, 224, taps=taps, **kw)
    t = tok.cpu().view(n, L, -1)
    return t[:, 0], t[:, 1:], taps


EXTRAS = dict(sim_cfg={}, outlier_cfg={'top_k': 30})


@pytest.mark.parametrize('precision,tol', [('fp32', 2e-4), ('bf16', 6e-2)])
def test_vit_tiny_all_model_types(gold, precision, tol):
    """ViT-tiny-16 (4 layers, width 128), 2 crops, extras ON, every custom_attn variant, vs the reference."""
    g = gold('vit_tiny')
    cfg, eng = _visual_engine('ViT-tiny-16', precision)
    for mt in ['Experimental', 'SCLIP', 'ClearCLIP', 'SFP', 'vanilla', 'SegEarth', 'MaskCLIP']:
        cls, tok, taps = _encode(eng, 2, model_type=mt, **EXTRAS)
        if precision == 'fp32' and mt == 'Experimental':
            assert np.array_equal(taps['outlier_idx'].cpu().numpy(), g['tap_outlier_idx'])
            assert np.abs(taps['simmap'].cpu().numpy()[:, ::7] - g['tap_simmap']).max() < 1e-5
            assert np.abs(taps['ln_pre'].cpu().view(2, 197, -1).numpy()[:, ::9] - g['tap_ln_pre']).max() < 1e-4
        e_tok = np.abs(tok.numpy() - g[f'{mt}_tokens']).max()
        e_cls = np.abs(cls.numpy() - g[f'{mt}_cls']).max()
        print(f'[vit_tiny {precision} {mt}] max|dtok|={e_tok:.2e} max|dcls|={e_cls:.2e}')
        if precision == 'fp32':
            assert e_tok < tol and e_cls < tol, (mt, e_tok, e_cls)
        else:
            # bf16: the top-k outlier selection is discrete -- a rounding-level change of the attention
            # statistics may swap two near-tied patches, which rewrites a handful of token rows.  Require
            # the tolerance on >= 97 % of the token rows and a loose bound on the rest.
            # A swapped patch itself differs by O(|token|) (suppressed in one run, kept in the other) and so do the
            # suppressed patches next to it (their replacement averages the non-outlier neighbours), so the rows
            # far outside the tolerance are bounded by the number of swaps: at most 2 swaps x 2 patches x 9 cells.
            row_err = np.abs(tok.numpy() - g[f'{mt}_tokens']).max(-1)
            swaps = 2
            if mt == 'Experimental':
                mine, ref = taps['outlier_idx'].cpu().numpy(), g['tap_outlier_idx']
                swaps = sum(len(set(a.tolist()) - set(b.tolist())) for a, b in zip(mine, ref))
                assert swaps <= 2
            far, out = int((row_err >= 10 * tol).sum()), float((row_err >= tol).mean())
            print(f'    outlier swaps vs fp32: {swaps}; rows outside tol: {out * 100:.2f}%, rows beyond 10 x tol: {far}')
            assert out <= 0.05 and far <= 18 * swaps and e_cls < tol, (mt, e_tok, e_cls)
    cls, tok, _ = _encode(eng, 2, model_type='Experimental')
    assert np.abs(tok.numpy() - g['plain_tokens']).max() < tol
    cls, tok, _ = _encode(eng, 2, model_type='Experimental', ignore_residual=False)
    assert np.abs(tok.numpy() - g['residual_tokens']).max() < tol


def test_vit_tiny_quickgelu(gold):
    from clip_decontamination_b200.engine import VisualEngine
    g = gold('vit_tiny')
    cfg = get_model_config('ViT-tiny-16')
    v = cfg['vision_cfg']
    eng = VisualEngine(_visual_sd(cfg), width=v['width'], layers=v['layers'], heads=v['heads'],
                       patch_size=16, image_size=224, embed_dim=cfg['embed_dim'], quick_gelu=True, precision='fp32')
    cls, tok, _ = _encode(eng, 2, model_type='Experimental')
    assert np.abs(tok.numpy() - g['quickgelu_tokens']).max() < 2e-4


@pytest.mark.parametrize('precision,tol', [('fp32', 3e-4), ('bf16', 8e-2)])
def test_vit_b16_crop(gold, precision, tol):
    """ViT-B/16, one 224 crop, extras ON: tokens (mean |.| 0.77) vs the reference."""
    g = gold('vit_b16_crop')
    cfg, eng = _visual_engine('ViT-B-16', precision)
    cls, tok, taps = _encode(eng, 1, model_type='Experimental', **EXTRAS)
    e = np.abs(tok.numpy() - g['Experimental_tokens'])
    print(f'[vit_b16 {precision}] tokens max|d|={e.max():.3e} mean|d|={e.mean():.3e}; '
          f'block10 max|d|={np.abs(taps["block10"].cpu().view(1,197,-1).numpy()[:, ::9] - g["tap_block10"]).max():.3e}')
    if precision == 'fp32':
        assert np.array_equal(taps['outlier_idx'].cpu().numpy(), g['tap_outlier_idx'])
    assert e.max() < tol
    cls, tok, _ = _encode(eng, 1, model_type='Experimental')
    assert np.abs(tok.numpy() - g['plain_tokens']).max() < tol


@pytest.mark.parametrize('name', ['jbu_one', 'jbu_stack'])
@pytest.mark.parametrize('precision,tol', [('fp32', 1e-4), ('bf16', 4e-2)])
def test_jbu_small(gold, name, precision, tol):
    """JBUOne / JBUStack with C=32 on a 224 crop, stage by stage, vs the reference."""
    from clip_decontamination_b200.engine import JBUEngine
    g = gold(f'{name}_c32')
    eng = JBUEngine(name, synthetic_jbu_state_dict(name, 32, 1), 32, precision)
    img = torch.from_numpy(synth.preprocess(synth.voronoi_scene(224, 224, 5))).cuda()
    wins = torch.tensor([(0, 0, 224, 224)], dtype=torch.int32).cuda()
    src = torch.from_numpy(g['source'])                                  # [1,32,14,14]
    feats = src[0].permute(1, 2, 0).reshape(196, 32).contiguous().to(eng.cdt).cuda()
    taps = {}
    out = eng.upsample(feats, 14, 14, img, wins, 224, 224, taps=taps)
    nchw = lambda t, h: t.float().cpu().view(1, h, h, -1).permute(0, 3, 1, 2)
    d2 = g['kernel0'].shape[1]
    k0 = taps['jbu_kernels'][0].float().cpu().view(1, 28, 28, -1)[..., :d2].permute(0, 3, 1, 2)
    print(f'[{name} {precision}] kernel0 max|d|={np.abs(k0.numpy() - g["kernel0"]).max():.3e}')
    assert np.abs(k0.numpy() - g['kernel0']).max() < (1e-5 if precision == "fp32" else 6e-3)
    errs = [np.abs(nchw(taps['jbu_stages'][0], 28).numpy() - g['stage0']).max(),
            np.abs(nchw(taps['jbu_stages'][1], 56).numpy() - g['stage1']).max(),
            np.abs(nchw(taps['jbu_stages'][2], 112).numpy()[:, :, ::2, ::2] - g['stage2']).max(),
            np.abs(nchw(taps['jbu_stages'][3], 224).numpy()[:, :, ::4, ::4] - g['stage3']).max()]
    o = nchw(out, 224)
    e = np.abs(o.numpy()[:, :, ::4, ::4] - g['out']).max()
    print(f'[{name} {precision}] stage errs={["%.2e" % x for x in errs]} out={e:.2e}')
    assert max(errs) < tol and e < tol
    if precision == 'fp32':                                              # checksum over the FULL output
        assert abs(float(o.double().sum()) - float(g['out_sum'])) < 1e-3 * abs(float(g['out_sqsum'])) ** 0.5


def test_jbu_stack_real_checkpoint(gold):
    """The only learned weights shipped with the reference (clip_jbu_stack_cocostuff.ckpt, C=512)."""
    from clip_decontamination_b200.engine import JBUEngine
    g = gold('jbu_stack_real')
    sd = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('w.')}
    img = torch.from_numpy(synth.preprocess(synth.voronoi_scene(224, 224, 5))).cuda()
    wins = torch.tensor([(0, 0, 224, 224)], dtype=torch.int32).cuda()
    src = torch.randn(1, 512, 14, 14, generator=torch.Generator().manual_seed(7))
    for precision, tol in (('fp32', 1e-4), ('bf16', 4e-2)):
        eng = JBUEngine('jbu_stack', sd, 512, precision)
        feats = src[0].permute(1, 2, 0).reshape(196, 512).contiguous().to(eng.cdt).cuda()
        out = eng.upsample(feats, 14, 14, img, wins, 224, 224)
        o = out.float().cpu().view(1, 224, 224, 512).permute(0, 3, 1, 2)
        e = np.abs(o.numpy()[:, ::16, ::4, ::4] - g['out']).max()
        print(f'[jbu_stack_real {precision}] max|d|={e:.3e}')
        assert e < tol


def _seg_engine(model, cls, precision, g, upsampler=None, extras=True, thd=0.1, bg=5):
    from clip_decontamination_b200.engine import JBUEngine, SegEngine
    cfg, vis = _visual_engine(model, precision)
    up = None
    if upsampler:
        up = JBUEngine(upsampler, synthetic_jbu_state_dict(upsampler, cfg['embed_dim'], 1), cfg['embed_dim'], precision)
    return SegEngine(vis, torch.from_numpy(g['query_features']), g['query_idx'].tolist(), prob_thd=thd, bg_idx=bg,
                     global_debias_factor=0.2 if extras else 0.0, upsampler=up,
                     sim_cfg={} if extras else None, outlier_cfg={'top_k': 30} if extras else None)


def _check_seg(tag, seg, g, precision, logit_tol):
    H, W, thd, bg, seed = [int(g['meta'][0]), int(g['meta'][1]), float(g['meta'][2]), int(g['meta'][3]), int(g['meta'][4])]
    img = torch.from_numpy(synth.preprocess(synth.voronoi_scene(H, W, seed))).cuda()
    labels, _, avg = seg.segment(img, want_logits=True)
    torch.cuda.synchronize()
    e = np.abs(avg.cpu().numpy()[:, ::4, ::4] - g['logits_sub']).max()
    lab = labels.cpu().numpy()
    agree = (lab == g['labels']).mean()
    margin = g['margin'].astype(np.float32)
    safe = margin > 2 * max(e, 1e-7)
    agree_safe = (lab == g['labels'])[safe].mean() if safe.any() else 1.0
    print(f'[{tag} {precision}] max|dlogit|={e:.3e} label agreement={agree * 100:.3f}% '
          f'(on {safe.mean() * 100:.1f}% pixels with margin>2*err: {agree_safe * 100:.4f}%) '
          f'hist={np.bincount(lab.ravel(), minlength=seg.K).tolist()} ref_hist={g["hist"].tolist()}')
    assert e < logit_tol
    assert agree_safe >= 0.999
    if precision == 'fp32':
        assert agree >= 0.999
    return e, agree


@pytest.mark.parametrize('precision,tol', [('fp32', 1e-4), ('bf16', 1e-2)])
def test_seg_tiny_jbu(gold, precision, tol):
    """300x260 image (4 windows incl. snapped ones), ViT-tiny + jbu_one + extras, full predict path."""
    g = gold('seg_tiny_jbu')
    seg = _seg_engine('ViT-tiny-16', 'potsdam', precision, g, upsampler='jbu_one')
    _check_seg('seg_tiny_jbu', seg, g, precision, tol)


def test_similarity_map_buffers_survive_layout_changes(gold):
    """One engine, sequences of different token counts: the transposed similarity map has two layouts (208 / 272 key
    columns) whose zero CLS row / column are only written at allocation.  Whole-image calls with L = 226 (272 columns), 290
    (plain map) and 170 (208 columns), then the sliding 512x512 scene (L = 197) must reproduce a fresh engine's result."""
    g = gold('seg_potsdam_noup')
    seg = _seg_engine('ViT-B-16', 'potsdam', 'bf16', g)
    crop = seg.crop
    for hw in ((240, 240), (272, 272), (208, 208)):        # L = 226 (272 columns), 290 (beyond both: plain map), 170 (208)
        seg.crop = 0
        lab = seg.segment(torch.from_numpy(synth.preprocess(synth.voronoi_scene(hw[0], hw[1], 5))).cuda())[0]
        assert lab.shape == hw
    seg.crop = crop
    e, _ = _check_seg('seg_potsdam_noup after whole-image calls', seg, g, 'bf16', 1e-2)
    fresh = _seg_engine('ViT-B-16', 'potsdam', 'bf16', g)
    e0, _ = _check_seg('seg_potsdam_noup fresh engine', fresh, g, 'bf16', 1e-2)
    assert e == e0                                         # bit-identical logits error: nothing stale was read


@pytest.mark.parametrize('precision,tol', [('fp32', 1e-4), ('bf16', 1e-2)])
def test_seg_potsdam_noup(gold, precision, tol):
    """BASELINE config 1 without the upsampler: 512x512, ViT-B/16, 16 crops, extras ON."""
    g = gold('seg_potsdam_noup')
    seg = _seg_engine('ViT-B-16', 'potsdam', precision, g)
    _check_seg('seg_potsdam_noup', seg, g, precision, tol)


@pytest.mark.parametrize('precision,tol', [('fp32', 1e-4), ('bf16', 1e-2)])
def test_seg_potsdam_jbu(gold, precision, tol):
    """BASELINE config 1/2: 512x512, ViT-B/16 + jbu_one, 16 crops, extras ON."""
    g = gold('seg_potsdam_jbu')
    seg = _seg_engine('ViT-B-16', 'potsdam', precision, g, upsampler='jbu_one')
    _check_seg('seg_potsdam_jbu', seg, g, precision, tol)


def test_seg_potsdam_jbu_basis_vs_literal(gold):
    """bf16 head in basis form (upsample token indicators, Gram contraction: cseg_basis_logits) against the
    literal form (upsample 512 channels, fused 1x1 conv + normalise: cseg_fixup_norm_sim) on the same crops:
    per-crop cosine logits agree to 5e-3 and both stay within the 1e-2 bar of the fp32 reference."""
    g = gold('seg_potsdam_jbu')
    seg = _seg_engine('ViT-B-16', 'potsdam', 'bf16', g, upsampler='jbu_one')
    assert seg.basis and seg.up.basis_ok(196, 224 * 224, seg.Q)
    H, W, seed = int(g['meta'][0]), int(g['meta'][1]), int(g['meta'][4])
    img = torch.from_numpy(synth.preprocess(synth.voronoi_scene(H, W, seed))).cuda()
    lb = seg.crop_logits(img)[0].clone()
    seg.basis = False
    ll = seg.crop_logits(img)[0].clone()
    torch.cuda.synchronize()
    d = (lb - ll).abs().max().item()
    print(f'[basis vs literal] max|dlogit| per crop = {d:.3e}')
    assert torch.isfinite(lb).all() and d < 5e-3


@pytest.mark.parametrize('batch', [1, 2])
def test_jbu_shared_kernels_equal_per_crop(gold, batch):
    """JBU kernel generation shared across overlapping crops (image-level tensors + per-crop border frames,
    csrc/jbu_share.cuh) against the per-crop form: the same arithmetic per pixel, so the crop logits are identical."""
    g = gold('seg_potsdam_jbu')
    seg = _seg_engine('ViT-B-16', 'potsdam', 'bf16', g, upsampler='jbu_one')
    H = W = 512
    img = torch.cat([torch.from_numpy(synth.preprocess(synth.voronoi_scene(H, W, 2 + b))) for b in range(batch)], 1).cuda()
    assert seg.up.share_ok(seg._windows(H, W, batch)[1], batch * H * W, 224, 224, 0, 0)
    seg.share_kernels = True
    a = seg.crop_logits(img, batch=batch)[0].clone()
    seg.share_kernels = False
    b = seg.crop_logits(img, batch=batch)[0].clone()
    torch.cuda.synchronize()
    d = (a - b).abs().max().item()
    print(f'[shared vs per-crop JBU kernels, batch {batch}] max|dlogit| = {d:.3e}')
    assert torch.isfinite(a).all() and d <= 1e-6
    # literal (non-basis) head too
    seg.basis = False
    seg.share_kernels = True
    a = seg.crop_logits(img, batch=batch)[0].clone()
    seg.share_kernels = False
    b = seg.crop_logits(img, batch=batch)[0].clone()
    assert (a - b).abs().max().item() <= 1e-6
    # windows off the 16-pixel lattice (snapped last window of a 1300-wide image) take the per-crop path
    assert not seg.up.share_ok(seg._windows(1300, 1100)[1], 1300 * 1100, 224, 224, 0, 0)


@pytest.mark.parametrize('upsampler', [None, 'jbu_one'])
def test_batched_equals_per_image(gold, upsampler):
    """A batch [B,H,W,3] goes through every kernel as B x 16 crops (stacked canvas); crop-level work is independent,
    so the labels equal those of B single-image calls (slide_inference with a batched input, segmentor.py:413-449)."""
    g = gold('seg_potsdam_jbu')
    seg = _seg_engine('ViT-B-16', 'potsdam', 'bf16', g, upsampler=upsampler)
    B, H, W = 3, 512, 512
    u8 = torch.stack([torch.from_numpy(synth.voronoi_scene(H, W, 70 + b)) for b in range(B)]).cuda()
    single = torch.stack([seg.segment_u8(u8[b], use_graph=False).clone() for b in range(B)])
    batched = seg.segment_u8(u8, use_graph=False)
    graphed = seg.segment_u8(u8, use_graph=True).clone()
    torch.cuda.synchronize()
    same = (single == batched).float().mean().item()
    print(f'[batched vs single, up={upsampler}] label agreement {same * 100:.4f}%')
    assert same >= 0.9999
    assert torch.equal(batched, graphed)


@pytest.mark.parametrize('precision,tol', [('fp32', 1e-4), ('bf16', 1e-2)])
def test_seg_loveda_vitl(gold, precision, tol):
    """BASELINE config 3 shape: ViT-L/14 (L=257, 24 layers), no upsampler, 448x448 (9 crops), Q=9 -> K=7."""
    g = gold('seg_loveda_vitl')
    seg = _seg_engine('ViT-L-14', 'loveda', precision, g, thd=0.3, bg=0)
    _check_seg('seg_loveda_vitl', seg, g, precision, tol)
