"""BASELINE.json's full-size configurations through size-independent properties (the CPU oracle needs
minutes per tile at these sizes): determinism, independence from the crop chunking, CUDA-graph replay ==
eager launches, consistency of labels with the probability output and the threshold rule, additivity of the
IoU histograms."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from clip_decontamination_b200 import synth  # noqa: E402
from clip_decontamination_b200.open_clip.model_configs import get_model_config  # noqa: E402
from clip_decontamination_b200.open_clip.synthetic import synthetic_clip_state_dict, synthetic_jbu_state_dict  # noqa: E402

EXTRAS = dict(sim_cfg={}, outlier_cfg={'top_k': 30})


def _engine(model, qf, qidx, thd, bg, upsampler, jbu_chunk=16):
    from clip_decontamination_b200.engine import VisualEngine, JBUEngine, SegEngine
    cfg = get_model_config(model)
    v = cfg['vision_cfg']
    sd = synthetic_clip_state_dict(cfg, 0, text_tower=False)
    vis = VisualEngine({k[len('visual.'):]: t for k, t in sd.items() if k.startswith('visual.')}, width=v['width'],
                       layers=v['layers'], heads=v['heads'], patch_size=v['patch_size'], image_size=v['image_size'],
                       embed_dim=cfg['embed_dim'], precision='bf16')
    up = JBUEngine(upsampler, synthetic_jbu_state_dict(upsampler, cfg['embed_dim'], 1), cfg['embed_dim'], 'bf16') \
        if upsampler else None
    return SegEngine(vis, qf, qidx, prob_thd=thd, bg_idx=bg, global_debias_factor=0.2, upsampler=up,
                     jbu_chunk=jbu_chunk, **EXTRAS)


@pytest.mark.parametrize('tag,H,W,cls,thd,bg', [('isaid 896^2 (config 4)', 896, 896, 'isaid', 0.4, 0),
                                                 ('road 1024^2 (config 5)', 1024, 1024, 'roadval', 0.7, 0),
                                                 ('road 1300x1100 (snapped windows)', 1300, 1100, 'roadval', 0.7, 0)])
def test_vit_b16_jbu_full_size(gold, tag, H, W, cls, thd, bg):
    g = gold('bench_text')
    qf, qidx = torch.from_numpy(g[f'{cls}_query_features']), g[f'{cls}_query_idx'].tolist()
    u8 = torch.from_numpy(synth.voronoi_scene(H, W, 21)).cuda()
    img = torch.from_numpy(synth.preprocess(synth.voronoi_scene(H, W, 21))).cuda()
    eng = _engine('ViT-B-16', qf, qidx, thd, bg, 'jbu_one')
    lab1, probs, avg = eng.segment(img, want_probs=True, want_logits=True)
    lab1, probs, avg = lab1.clone(), probs.clone(), avg.clone()
    lab2, _, _ = eng.segment(img)
    assert torch.equal(lab1, lab2), 'not deterministic'
    # crops are independent: a different JBU chunking must give bit-identical logits and labels
    eng7 = _engine('ViT-B-16', qf, qidx, thd, bg, 'jbu_one', jbu_chunk=7)
    lab7, _, avg7 = eng7.segment(img, want_logits=True)
    assert torch.equal(avg, avg7) and torch.equal(lab1, lab7)
    # CUDA-graph replay of the uint8 path == eager launches
    labg = eng.segment_u8(u8).clone()
    labe = eng.segment_u8(u8, use_graph=False)
    assert torch.equal(labg, labe) and torch.equal(labg, lab1)
    # labels are consistent with the probabilities: argmax (lowest index on ties) and the prob_thd rule
    K = eng.K
    pmax, parg = probs.max(0)
    expect = torch.where(pmax < thd, torch.full_like(parg, bg), parg)
    assert (expect == lab1.long()).float().mean().item() > 0.9999     # == up to exact float ties
    assert lab1.max().item() < K and torch.isfinite(avg).all()
    assert abs(float(probs.sum(0).mean()) - 1.0) < 1e-3 or K != eng.Q
    # histogram additivity over quadrants
    from clip_decontamination_b200 import ops
    gt = torch.from_numpy(synth.synthetic_labels(H, W, K, 5)).cuda()
    full = torch.zeros((3, K), dtype=torch.int64, device='cuda')
    ops.iou_hist(lab1.view(-1), gt.view(-1), K, full)
    parts = torch.zeros_like(full)
    for ys in (slice(0, H // 2), slice(H // 2, H)):
        for xs in (slice(0, W // 2), slice(W // 2, W)):
            ops.iou_hist(lab1[ys, xs].contiguous().view(-1), gt[ys, xs].contiguous().view(-1), K, parts)
    assert torch.equal(full, parts)
    print(f'[{tag}] crops={len(eng._windows(H, W)[1])} label hist={torch.bincount(lab1.view(-1).long(), minlength=K).tolist()}')


def test_vit_l14_1024_config3(gold):
    """BASELINE config 3: ViT-L/14 (L=257, 24 layers), no upsampler, 1024^2 (81 crops), Q=9 -> K=7."""
    g = gold('seg_loveda_vitl')
    eng = _engine('ViT-L-14', torch.from_numpy(g['query_features']), g['query_idx'].tolist(), 0.3, 0, None)
    img = torch.from_numpy(synth.preprocess(synth.voronoi_scene(1024, 1024, 22))).cuda()
    lab1, _, avg = eng.segment(img, want_logits=True)
    lab1, avg = lab1.clone(), avg.clone()
    lab2, _, _ = eng.segment(img)
    assert torch.equal(lab1, lab2) and torch.isfinite(avg).all()
    assert len(eng._windows(1024, 1024)[1]) == 81 and lab1.max().item() < 7
    # the top-left 448x448 corner sees exactly the windows of a 448x448 image except along its right/bottom
    # 112-pixel band; in the interior [0,336)^2 the averaged logits must equal those of the small image
    small, _, avg_s = eng.segment(img[:, :448, :448].contiguous(), want_logits=True)
    assert (avg[:, :336, :336] - avg_s[:, :336, :336]).abs().max().item() < 1e-6
