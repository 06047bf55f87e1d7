"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference Python
(imported from /root/reference through oracle/ref_harness.py) and, in the same pass, pin the
oracle (oracle/clipseg_oracle.py) against it.  TEST INFRASTRUCTURE; container-only (needs
/root/reference).  Usage:  python -m oracle.gen_golden [group ...]

Groups: vit_tiny vit_b16 jbu_small jbu_real postproc text seg_noup seg_jbu seg_vitl
Every array is produced by reference code; the oracle's result on the same input must agree
within ``PIN_TOL`` or the script aborts (so a committed fixture implies a pinned oracle).
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh                      # noqa: E402
from oracle import clipseg_oracle as O                    # noqa: E402
from clip_decontamination_b200.open_clip.model_configs import get_model_config          # noqa: E402
from clip_decontamination_b200.open_clip.synthetic import (synthetic_clip_state_dict,   # noqa: E402
                                                           synthetic_jbu_state_dict)
from clip_decontamination_b200 import synth               # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
PIN_TOL = 2e-5
EXTRAS = dict(global_debias_factor=0.2, apply_outlier_suppression=True,
              outlier_suppression_cfg=dict(top_k=30), apply_similarity_enhancement=True,
              similarity_enhancement_cfg=dict(similarity_weight=1.0, temperature=1.0,
                                              add_self_similarity=True))


def _pin(name, ref, orc, tol=PIN_TOL):
    d = float((ref.float() - orc.float()).abs().max())
    print(f'  pin {name}: max|ref-oracle| = {d:.3e}')
    assert d <= tol, f'oracle disagrees with the reference on {name}: {d}'


def _save(name, **arrs):
    path = os.path.join(GOLD, name + '.npz')
    np.savez_compressed(path, **{k: (v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
                                 for k, v in arrs.items()})
    print(f'  wrote {path} ({os.path.getsize(path) / 1e3:.0f} kB)')


def _visual(sd):
    return {k[len('visual.'):]: v for k, v in sd.items() if k.startswith('visual.')}


def _two_crops(seed=5):
    img = torch.from_numpy(synth.preprocess(synth.voronoi_scene(224, 448, seed)))
    return torch.stack([img[:, :, :224], img[:, :, 224:]])


def _ref_visual(cfg, sd, extras=True):
    m = rh.build_ref_clip(cfg, sd)
    if extras:
        from similarity_enhancement import SimilarityEnhancementModule
        from outlier_suppression import OutlierSuppressionModule
        m.visual.similarity_enhancer = SimilarityEnhancementModule()
        m.visual.outlier_suppressor = OutlierSuppressionModule(top_k=30)
    return m


@torch.no_grad()
def vit_group(model_name, out_name, model_types):
    cfg = get_model_config(model_name)
    sd = synthetic_clip_state_dict(cfg, 0)
    v = cfg['vision_cfg']
    x = _two_crops() if model_name == 'ViT-tiny-16' else _two_crops()[:1]
    vis = _visual(sd)
    arrs = {}
    m = _ref_visual(cfg, sd, True)
    for mt in model_types:
        rc, rt = m.encode_image(x, mt, True, output_cls_token=True, apply_similarity_enhancement=True)
        taps = {}
        oc, ot = O.vit_dense_forward(vis, x, layers=v['layers'], heads=v['heads'], patch=v['patch_size'],
                                     model_type=mt, sim_cfg={}, outlier_cfg={'top_k': 30}, taps=taps)
        _pin(f'{model_name}/{mt}/cls', rc, oc)
        _pin(f'{model_name}/{mt}/tokens', rt, ot)
        arrs[f'{mt}_cls'], arrs[f'{mt}_tokens'] = rc, rt
        if mt == 'Experimental':
            # intermediates come from the oracle (the reference exposes no taps); they are only
            # used to localise drift, the pinned outputs are cls/tokens above.
            arrs['tap_outlier_idx'] = taps['outlier_idx']
            arrs['tap_stats_cls'] = taps['attn_stats_cls']
            arrs['tap_stats_diag'] = taps['attn_stats_diag']
            arrs['tap_simmap'] = taps['simmap'][:, ::7, :]
            arrs['tap_final_attn'] = taps['final_attn'][:, ::5]
            arrs['tap_suppressed'] = taps['suppressed'][:, ::5]
            arrs['tap_ln_pre'] = taps['ln_pre'][:, ::9]
            arrs['tap_block0'] = taps['block0'][:, ::9]
            arrs[f'tap_block{v["layers"] - 2}'] = taps[f'block{v["layers"] - 2}'][:, ::9]
    m = _ref_visual(cfg, sd, False)
    rc, rt = m.encode_image(x, 'Experimental', True, output_cls_token=True)
    oc, ot = O.vit_dense_forward(vis, x, layers=v['layers'], heads=v['heads'], patch=v['patch_size'])
    _pin(f'{model_name}/plain/tokens', rt, ot)
    arrs['plain_cls'], arrs['plain_tokens'] = rc, rt
    rc, rt = m.encode_image(x, 'Experimental', False, output_cls_token=True)
    oc, ot = O.vit_dense_forward(vis, x, layers=v['layers'], heads=v['heads'], patch=v['patch_size'],
                                 ignore_residual=False)
    _pin(f'{model_name}/residual/tokens', rt, ot)
    arrs['residual_cls'], arrs['residual_tokens'] = rc, rt
    if model_name == 'ViT-tiny-16':
        cfgq = dict(cfg, quick_gelu=True)
        mq = rh.build_ref_clip(cfgq, sd)
        rc, rt = mq.encode_image(x, 'Experimental', True, output_cls_token=True)
        oc, ot = O.vit_dense_forward(vis, x, layers=v['layers'], heads=v['heads'], patch=v['patch_size'],
                                     quick_gelu=True)
        _pin(f'{model_name}/quickgelu/tokens', rt, ot)
        arrs['quickgelu_cls'], arrs['quickgelu_tokens'] = rc, rt
    _save(out_name, **arrs)


@torch.no_grad()
def jbu_small():
    rh.install()
    from simfeatup_dev.upsamplers import get_upsampler
    img = torch.from_numpy(synth.preprocess(synth.voronoi_scene(224, 224, 5)))[None]
    for name, C in (('jbu_one', 32), ('jbu_stack', 32)):
        sd = synthetic_jbu_state_dict(name, C, 1)
        up = get_upsampler(name, C).eval()
        up.load_state_dict(sd, strict=True)
        src = torch.randn(1, C, 14, 14, generator=torch.Generator().manual_seed(7))
        # reference stage outputs: call the stage method the way JBUOne/JBUStack.forward does
        mods = [up.up] * 4 if name == 'jbu_one' else [up.up1, up.up2, up.up3, up.up4]
        s, stages = src, []
        for mod in mods:
            s = up.upsample(s, img, mod)
            stages.append(s)
        out = up(src, img)
        taps = {}
        o = O.jbu_upsample(sd, name, src, img, taps)
        _pin(f'{name}/out', out, o)
        for i in range(4):
            _pin(f'{name}/stage{i}', stages[i], taps['jbu_stages'][i])
        _save(f'{name}_c32', source=src, stage0=stages[0], stage1=stages[1], stage2=stages[2][:, :, ::2, ::2],
              stage3=stages[3][:, :, ::4, ::4], out=out[:, :, ::4, ::4],
              out_sum=np.float64(out.double().sum()), out_sqsum=np.float64(out.double().square().sum()),
              kernel0=taps['jbu_kernels'][0], kernel3=taps['jbu_kernels'][3][:, :, ::8, ::8])


@torch.no_grad()
def jbu_real():
    rh.install()
    from simfeatup_dev.upsamplers import get_upsampler
    ck = torch.load(os.path.join(rh.REF, 'simfeatup_dev/weights/clip_jbu_stack_cocostuff.ckpt'),
                    weights_only=False, map_location='cpu')['state_dict']
    sd = {k[10:]: v.float() for k, v in ck.items()}                      # segmentor.py:282
    up = get_upsampler('jbu_stack', 512).eval()
    up.load_state_dict(sd, strict=True)
    img = torch.from_numpy(synth.preprocess(synth.voronoi_scene(224, 224, 5)))[None]
    src = torch.randn(1, 512, 14, 14, generator=torch.Generator().manual_seed(7))
    out = up(src, img)
    o = O.jbu_upsample(sd, 'jbu_stack', src, img)
    _pin('jbu_stack_real/out', out, o)
    arrs = {'w.' + k: v for k, v in sd.items()}
    _save('jbu_stack_real', out=out[:, ::16, ::4, ::4], out_sum=np.float64(out.double().sum()),
          out_sqsum=np.float64(out.double().square().sum()), **arrs)


@torch.no_grad()
def postproc():
    cfg = get_model_config('ViT-tiny-16')
    sd = synthetic_clip_state_dict(cfg, 0)
    g = torch.Generator().manual_seed(11)
    cases = {}
    for tag, cls, H, W, thd, bg, stride, crop in (('potsdam', 'potsdam', 100, 90, 0.3, 5, 32, 64),
                                                  ('loveda', 'loveda', 64, 150, 0.25, 0, 32, 64),
                                                  ('road', 'roadval', 131, 97, 0.7, 0, 24, 48)):
        seg = rh.build_ref_segmentor(cfg, sd, os.path.join(ROOT, 'configs', f'cls_{cls}.txt'),
                                     model_type='Experimental', prob_thd=thd, bg_idx=bg)
        Q = seg.num_queries
        wins = O.slide_windows(H, W, stride, crop)
        crop_logits = torch.randn(len(wins), Q, min(crop, H), min(crop, W), generator=g) * 0.03
        # reference accumulation loop (segmentor.py:413-449) on given crop logits
        it = iter(crop_logits)
        seg.forward_feature = lambda c, **kw: next(it)[None]
        lg = seg.forward_slide(torch.zeros(1, 3, H, W), [dict(ori_shape=(H, W))], stride, crop)
        pred = seg.postprocess_result(lg.clone(), None)
        avg, pr, opred = O.postprocess_from_crop_logits(crop_logits, wins, H, W, seg.query_idx.tolist(),
                                                        50, thd, bg)
        _pin(f'postproc/{tag}/avg', lg[0], avg, 1e-7)
        assert torch.equal(pred, opred), 'oracle labels differ from the reference'
        cases[f'{tag}_crop_logits'] = crop_logits.half()     # stored as fp16; tests recompute from these
        cases[f'{tag}_meta'] = np.array([H, W, thd, bg, stride, crop], dtype=np.float64)
        # labels for the fp16-rounded logits (what the test feeds), from the reference
        cl16 = crop_logits.half().float()
        it = iter(cl16)
        lg = seg.forward_slide(torch.zeros(1, 3, H, W), [dict(ori_shape=(H, W))], stride, crop)
        pred = seg.postprocess_result(lg.clone(), None)
        cases[f'{tag}_labels'] = pred[0].to(torch.uint8)
        cases[f'{tag}_query_idx'] = seg.query_idx
    _save('postproc', **cases)


@torch.no_grad()
def text():
    cfg = get_model_config('ViT-tiny-16')
    sd = synthetic_clip_state_dict(cfg, 0)
    rh.install()
    from open_clip import tokenizer
    from prompts.imagenet_template import openai_imagenet_template
    words = ['road', 'parking lot', 'low vegetation']
    prompts = [t(w) for w in words for t in openai_imagenet_template[:5]]
    toks = tokenizer.tokenize(prompts)
    m = rh.build_ref_clip(cfg, sd)
    feats = m.encode_text(toks)
    seg = rh.build_ref_segmentor(cfg, sd, os.path.join(ROOT, 'configs', 'cls_potsdam.txt'),
                                 model_type='Experimental')
    _save('text_tiny', prompts=np.array(prompts), tokens=toks, feats=feats,
          potsdam_query_features=seg.query_features, n_templates=np.int64(len(openai_imagenet_template)),
          template_probe=np.array([t('X') for t in openai_imagenet_template]))


@torch.no_grad()
def seg_group(out_name, model_name, cls, H, W, thd, bg, upsampler=None, extras=True, scene_seed=2):
    cfg = get_model_config(model_name)
    sd = synthetic_clip_state_dict(cfg, 0)
    v = cfg['vision_cfg']
    kw = dict(EXTRAS) if extras else {}
    up = None
    if upsampler:
        up = (upsampler, synthetic_jbu_state_dict(upsampler, cfg['embed_dim'], 1))
    t0 = time.time()
    seg = rh.build_ref_segmentor(cfg, sd, os.path.join(ROOT, 'configs', f'cls_{cls}.txt'),
                                 model_type='Experimental', prob_thd=thd, bg_idx=bg, upsampler=up, **kw)
    if 'L' in model_name:
        pass
    print(f'  ref init {time.time() - t0:.1f}s')
    img = torch.from_numpy(synth.preprocess(synth.voronoi_scene(H, W, scene_seed)))[None]
    t0 = time.time()
    lg = seg.forward_slide(img, [dict(ori_shape=(H, W))], 112, 224)
    pred = seg.postprocess_result(lg.clone(), None)
    t_ref = time.time() - t0
    print(f'  reference forward_slide+postprocess: {t_ref:.1f}s')
    orc = O.SegOracle(_visual(sd), seg.query_features, seg.query_idx.tolist(), layers=v['layers'],
                      heads=v['heads'], patch=v['patch_size'], prob_thd=thd, bg_idx=bg,
                      global_debias_factor=0.2 if extras else 0.0, upsampler=up,
                      sim_cfg={} if extras else None, outlier_cfg={'top_k': 30} if extras else None)
    olg = orc.forward_slide(img)
    _, opred = orc.postprocess(olg[0])
    _pin(f'{out_name}/logits', lg, olg, 1e-5)
    agree = float((pred == opred).float().mean())
    print(f'  oracle label agreement with the reference: {agree * 100:.4f}%')
    assert agree >= 0.9999
    s = torch.sort(lg[0], dim=0, descending=True)[0]
    margin = (s[0] - s[1])
    _save(out_name, query_features=seg.query_features, query_idx=seg.query_idx,
          logits_sub=lg[0][:, ::4, ::4], labels=pred[0].to(torch.uint8), margin=margin.half(),
          meta=np.array([H, W, thd, bg, scene_seed, t_ref], dtype=np.float64),
          hist=torch.bincount(pred.flatten(), minlength=seg.num_classes))


MARGIN_Q = 2e-4          # top-2 logit margins of the full-size groups are stored as uint8 multiples of this


@torch.no_grad()
def seg_full(out_name, model_name, cls, H, W, thd, bg, upsampler=None, extras=True, scene_seed=2,
             crop=224, stride=112, ori_shape=None, pin=True, sub=4):
    """Full-size BASELINE configurations through the UNMODIFIED reference (predict's two branches,
    segmentor.py:468-473): slide (crop > 0, optional resize to ori_shape) or whole image (crop <= 0)."""
    cfg = get_model_config(model_name)
    sd = synthetic_clip_state_dict(cfg, 0)
    v = cfg['vision_cfg']
    kw = dict(EXTRAS) if extras else {}
    up = (upsampler, synthetic_jbu_state_dict(upsampler, cfg['embed_dim'], 1)) if upsampler else None
    seg = rh.build_ref_segmentor(cfg, sd, os.path.join(ROOT, 'configs', f'cls_{cls}.txt'),
                                 model_type='Experimental', prob_thd=thd, bg_idx=bg, upsampler=up, **kw)
    img = torch.from_numpy(synth.preprocess(synth.voronoi_scene(H, W, scene_seed)))[None]
    ori = (H, W) if ori_shape is None else tuple(ori_shape)
    t0 = time.time()
    if crop > 0:
        lg = seg.forward_slide(img, [dict(ori_shape=ori)], stride, crop)
    else:
        lg = seg.forward_feature(img, ori)
    pred = seg.postprocess_result(lg.clone(), None)
    t_ref = time.time() - t0
    print(f'  reference: {t_ref:.1f}s  label hist {torch.bincount(pred.flatten(), minlength=seg.num_classes).tolist()}')
    if pin:
        orc = O.SegOracle(_visual(sd), seg.query_features, seg.query_idx.tolist(), layers=v['layers'],
                          heads=v['heads'], patch=v['patch_size'], prob_thd=thd, bg_idx=bg, slide_stride=stride,
                          slide_crop=crop, global_debias_factor=0.2 if extras else 0.0, upsampler=up,
                          sim_cfg={} if extras else None, outlier_cfg={'top_k': 30} if extras else None)
        olg = orc.forward_slide(img, ori) if crop > 0 else orc.forward_feature(img, ori)
        _, opred = orc.postprocess(olg[0])
        _pin(f'{out_name}/logits', lg, olg, 1e-5)
        agree = float((pred == opred).float().mean())
        print(f'  oracle label agreement with the reference: {agree * 100:.4f}%')
        assert agree >= 0.9999
    s = torch.sort(lg[0], dim=0, descending=True)[0]
    margin = (s[0] - s[1])
    # with random-init text embeddings the configured prob_thd often sends every pixel to bg_idx: also store the
    # reference's labels with the threshold off, so that label agreement at full size is not vacuous
    seg.prob_thd = 0.0
    pred0 = seg.postprocess_result(lg.clone(), None)
    seg.prob_thd = thd
    print(f'  label hist without threshold {torch.bincount(pred0.flatten(), minlength=seg.num_classes).tolist()}')
    _save(out_name, query_features=seg.query_features, query_idx=seg.query_idx,
          logits_sub=lg[0][:, ::sub, ::sub].half() if sub > 4 else lg[0][:, ::sub, ::sub],
          labels=pred[0].to(torch.uint8), labels_nothd=pred0[0].to(torch.uint8),
          margin_u8=(margin / MARGIN_Q).clamp(0, 255).to(torch.uint8),
          meta=np.array([H, W, thd, bg, scene_seed, t_ref, crop, stride, ori[0], ori[1], sub, MARGIN_Q,
                         1.0 if extras else 0.0], dtype=np.float64),
          hist=torch.bincount(pred.flatten(), minlength=seg.num_classes))


@torch.no_grad()
def ref_bf16_yardstick():
    """The reference's OWN reduced-precision noise floor: the unmodified reference run in bf16
    (weights via convert_weights_to_lp, bf16 input, bf16 upsampler; the adaptive-conv shim accumulates its
    121 taps in fp32 as a device kernel would) against the fp32 goldens of the same scenes."""
    cfg = get_model_config('ViT-B-16')
    sd = synthetic_clip_state_dict(cfg, 0)
    out = {}
    for tag, gold_name, cls, upn in (('potsdam_noup', 'seg_potsdam_noup', 'potsdam', None),
                                     ('potsdam_jbu', 'seg_potsdam_jbu', 'potsdam', 'jbu_one'),
                                     ('vaihingen_jbu', 'seg_vaihingen_jbu', 'vaihingen', 'jbu_one')):
        path = os.path.join(GOLD, gold_name + '.npz')
        if not os.path.exists(path):
            print(f'  skip {tag}: {gold_name}.npz missing')
            continue
        g = np.load(path)
        H, W, thd, bg, seed = int(g['meta'][0]), int(g['meta'][1]), float(g['meta'][2]), int(g['meta'][3]), int(g['meta'][4])
        up = (upn, synthetic_jbu_state_dict(upn, cfg['embed_dim'], 1)) if upn else None
        seg = rh.build_ref_segmentor(cfg, sd, os.path.join(ROOT, 'configs', f'cls_{cls}.txt'), precision='bf16',
                                     model_type='Experimental', prob_thd=thd, bg_idx=bg, upsampler=up, **EXTRAS)
        if up is not None:
            seg.upsampler.bfloat16()
        img = torch.from_numpy(synth.preprocess(synth.voronoi_scene(H, W, seed)))[None].bfloat16()
        t0 = time.time()
        lg = seg.forward_slide(img, [dict(ori_shape=(H, W))], 112, 224)
        pred = seg.postprocess_result(lg.clone(), None)[0].to(torch.uint8)
        lab32 = torch.from_numpy(g['labels'])
        sub = lg[0][:, ::4, ::4].float()
        dl = float((sub - torch.from_numpy(g['logits_sub']).float()).abs().max())
        agree = float((pred == lab32).float().mean())
        print(f'  {tag}: reference bf16 vs reference fp32: label agreement {agree * 100:.3f}%, '
              f'max|dlogit| {dl:.2e} ({time.time() - t0:.0f}s)')
        out[f'{tag}_agree'] = np.float64(agree)
        out[f'{tag}_max_dlogit'] = np.float64(dl)
        out[f'{tag}_labels_bf16'] = pred
    _save('ref_bf16_yardstick', **out)


@torch.no_grad()
def bench_text():
    """prompt-ensembled class text embeddings (segmentor.py:157-174) of the benchmark configs, produced by
    the reference's tokenizer + text tower on the synthetic ViT-B/16 weights."""
    cfg = get_model_config('ViT-B-16')
    sd = synthetic_clip_state_dict(cfg, 0)
    arrs = {}
    for cls in ('vaihingen', 'potsdam', 'isaid', 'roadval'):
        seg = rh.build_ref_segmentor(cfg, sd, os.path.join(ROOT, 'configs', f'cls_{cls}.txt'), model_type='Experimental')
        arrs[f'{cls}_query_features'] = seg.query_features
        arrs[f'{cls}_query_idx'] = seg.query_idx
    _save('bench_text', **arrs)


GROUPS = {
    'bench_text': bench_text,
    'vit_tiny': lambda: vit_group('ViT-tiny-16', 'vit_tiny',
                                  ['Experimental', 'SCLIP', 'ClearCLIP', 'SFP', 'vanilla', 'SegEarth', 'MaskCLIP']),
    'vit_b16': lambda: vit_group('ViT-B-16', 'vit_b16_crop', ['Experimental']),
    'jbu_small': jbu_small,
    'jbu_real': jbu_real,
    'postproc': postproc,
    'text': text,
    'seg_tiny': lambda: seg_group('seg_tiny_jbu', 'ViT-tiny-16', 'potsdam', 300, 260, 0.1, 5, upsampler='jbu_one'),
    'seg_noup': lambda: seg_group('seg_potsdam_noup', 'ViT-B-16', 'potsdam', 512, 512, 0.1, 5),
    'seg_jbu': lambda: seg_group('seg_potsdam_jbu', 'ViT-B-16', 'potsdam', 512, 512, 0.1, 5, upsampler='jbu_one'),
    'seg_vitl': lambda: seg_group('seg_loveda_vitl', 'ViT-L-14', 'loveda', 448, 448, 0.3, 0),
    # ---- BASELINE.json configs at their real size + the branches of predict (round 2) ----
    'seg_vaihingen': lambda: seg_full('seg_vaihingen_jbu', 'ViT-B-16', 'vaihingen', 512, 512, 0.1, 5, 'jbu_one'),
    'seg_extras_off': lambda: seg_full('seg_potsdam_jbu_extras_off', 'ViT-B-16', 'potsdam', 512, 512, 0.1, 5, 'jbu_one',
                                       extras=False),
    'seg_whole_jbu': lambda: seg_full('seg_whole_240_jbu', 'ViT-B-16', 'potsdam', 240, 240, 0.1, 5, 'jbu_one', crop=0),
    'seg_whole_noup': lambda: seg_full('seg_whole_250_noup', 'ViT-B-16', 'potsdam', 250, 250, 0.1, 5, None, crop=0),
    # whole image beyond 320 tokens: the long-sequence attention path (L = 626 / 577)
    'seg_whole_long_noup': lambda: seg_full('seg_whole_400_noup', 'ViT-B-16', 'potsdam', 400, 400, 0.1, 5, None, crop=0),
    'seg_whole_long_jbu': lambda: seg_full('seg_whole_384_jbu', 'ViT-B-16', 'vaihingen', 384, 384, 0.1, 5, 'jbu_one', crop=0),
    'seg_small_side': lambda: seg_full('seg_small_200_jbu', 'ViT-B-16', 'vaihingen', 200, 200, 0.1, 5, 'jbu_one'),
    'seg_nonsquare_off': lambda: seg_full('seg_nonsquare_200x180_off', 'ViT-B-16', 'potsdam', 200, 180, 0.1, 5, None,
                                          extras=False),
    'seg_road_resize': lambda: seg_full('seg_road_448to1024', 'ViT-B-16', 'roadval', 448, 448, 0.7, 0, 'jbu_one',
                                        ori_shape=(1024, 1024), sub=8),
    'seg_vitl_1024': lambda: seg_full('seg_loveda_vitl_1024', 'ViT-L-14', 'loveda', 1024, 1024, 0.3, 0, None, sub=8),
    'seg_isaid': lambda: seg_full('seg_isaid_896', 'ViT-B-16', 'isaid', 896, 896, 0.4, 0, 'jbu_one', sub=8),
    'seg_road': lambda: seg_full('seg_road_1024', 'ViT-B-16', 'roadval', 1024, 1024, 0.7, 0, 'jbu_one', sub=8),
    'seg_road_snapped': lambda: seg_full('seg_road_1300x1100', 'ViT-B-16', 'roadval', 1300, 1100, 0.7, 0, 'jbu_one',
                                         sub=8, pin=False),
    'ref_bf16': ref_bf16_yardstick,
}

if __name__ == '__main__':
    os.makedirs(GOLD, exist_ok=True)
    names = sys.argv[1:] or list(GROUPS)
    for n in names:
        print(f'[{n}]')
        t0 = time.time()
        GROUPS[n]()
        print(f'  done in {time.time() - t0:.1f}s')
