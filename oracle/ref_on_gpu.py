"""The UNMODIFIED reference on the GPU box (SURVEY.md §8d "same-box comparators"), from the files staged into
``baseline/_ref`` by ``oracle/stage_ref.py``.  MEASUREMENT INFRASTRUCTURE -- not imported by the product path.

  (a) reference, CPU fp32, all host cores: one full Vaihingen-shaped 512x512 tile through forward_slide +
      postprocess_result (16 crops + accumulate + post-process), timed once  -> s/tile, MP/s
  (b) reference, eager PyTorch on the B200 the way the reference itself runs (fp16 weights, crop at a time,
      cuda autocast upsampler): steady-state s/tile, MP/s
  (c) agreement table on the golden scene (tests/golden/seg_vaihingen_jbu.npz = reference fp32 on the CPU):
      reference-fp16-on-B200 vs reference-fp32 (the reference's own reduced-precision noise floor) beside
      ours-bf16 vs reference-fp32, raw and on the pixels with margin > 2 x measured logit error.

    python -m oracle.ref_on_gpu [--skip-cpu] [--out gpurun_out/ref_on_gpu.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh                                                     # noqa: E402
from clip_decontamination_b200 import synth                                             # noqa: E402
from clip_decontamination_b200.open_clip.model_configs import get_model_config          # noqa: E402
from clip_decontamination_b200.open_clip.synthetic import (synthetic_clip_state_dict,   # noqa: E402
                                                           synthetic_jbu_state_dict)

EXTRAS = dict(global_debias_factor=0.2, apply_outlier_suppression=True, outlier_suppression_cfg=dict(top_k=30),
              apply_similarity_enhancement=True,
              similarity_enhancement_cfg=dict(similarity_weight=1.0, temperature=1.0, add_self_similarity=True))


def agreement(lab, logits_sub, g, sub=4):
    ref_lab, mq = g['labels'], float(g['meta'][11])
    e = float(np.abs(logits_sub - g['logits_sub']).max())
    margin = g['margin_u8'].astype(np.float32) * mq
    safe = margin > 2 * e + mq
    return dict(max_dlogit=e, raw_agreement=float((lab == ref_lab).mean()),
                safe_fraction=float(safe.mean()),
                safe_agreement=float((lab == ref_lab)[safe].mean()) if safe.any() else 1.0)


@torch.no_grad()
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--skip-cpu', action='store_true')
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'ref_on_gpu.json'))
    ap.add_argument('--tiles', type=int, default=3)
    args = ap.parse_args()
    if not rh.available():
        raise SystemExit('reference not staged: run `python -m oracle.stage_ref` in the build container first')
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'seg_vaihingen_jbu.npz'))
    H, W, thd, bg, seed = int(g['meta'][0]), int(g['meta'][1]), float(g['meta'][2]), int(g['meta'][3]), int(g['meta'][4])
    cfg = get_model_config('ViT-B-16')
    sd = synthetic_clip_state_dict(cfg, 0)
    up = ('jbu_one', synthetic_jbu_state_dict('jbu_one', cfg['embed_dim'], 1))
    name_path = os.path.join(ROOT, 'configs', 'cls_vaihingen.txt')
    img = torch.from_numpy(synth.preprocess(synth.voronoi_scene(H, W, seed)))[None]
    res = dict(workload='Vaihingen-shaped 512x512 tile, ViT-B/16 + jbu_one, Q=K=6, 16 crops, extras ON',
               reference=rh.REF, torch=torch.__version__, gpu=torch.cuda.get_device_name(0))

    if not args.skip_cpu:                                            # (a) CPU fp32, all cores, one full tile
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        seg = rh.build_ref_segmentor(cfg, sd, name_path, model_type='Experimental', prob_thd=thd, bg_idx=bg,
                                     upsampler=up, **EXTRAS)
        t0 = time.time()
        lg = seg.forward_slide(img, [dict(ori_shape=(H, W))], 112, 224)
        pred = seg.postprocess_result(lg.clone(), None)
        dt = time.time() - t0
        a = agreement(pred[0].numpy().astype(np.uint8), lg[0][:, ::4, ::4].numpy(), g)
        res['reference_cpu_fp32'] = dict(seconds_per_tile=dt, mp_per_s=H * W / 1e6 / dt, cores=threads,
                                         threads=torch.get_num_threads(), vs_golden=a)
        print('[ref cpu fp32]', json.dumps(res['reference_cpu_fp32']))

    # (b) the reference's own GPU path: fp16, eager, crop at a time
    seg = rh.build_ref_segmentor_cuda(cfg, sd, name_path, model_type='Experimental', prob_thd=thd, bg_idx=bg,
                                      upsampler=up, **EXTRAS)
    x = img.cuda()
    pred = seg.predict(x, None)                                      # warm-up (cudnn / cublas handles)
    torch.cuda.synchronize()
    ts = []
    for _ in range(args.tiles):
        t0 = time.time()
        pred = seg.predict(x, None)
        torch.cuda.synchronize()
        ts.append(time.time() - t0)
    lg = seg.forward_slide(x.half(), [dict(ori_shape=(H, W))], 112, 224)
    a_ref16 = agreement(pred[0].cpu().numpy().astype(np.uint8), lg[0][:, ::4, ::4].float().cpu().numpy(), g)
    res['reference_b200_fp16_eager'] = dict(seconds_per_tile=float(np.median(ts)), mp_per_s=H * W / 1e6 / float(np.median(ts)),
                                            tiles_timed=len(ts), vs_golden=a_ref16)
    print('[ref b200 fp16 eager]', json.dumps(res['reference_b200_fp16_eager']))

    # (c) ours, bf16, same scene
    from clip_decontamination_b200.open_clip import create_model
    from clip_decontamination_b200.segmentor import SegmentorEx
    net = create_model('ViT-B/16', pretrained=None, precision='fp16')
    ours = SegmentorEx(clip_type='CLIP', vit_type='ViT-B/16', model_type='Experimental', name_path=name_path,
                       prob_thd=thd, bg_idx=bg, apply_sim_feat_up=True, sim_feat_up_cfg=dict(model_name='jbu_one', model_path=None),
                       net=net, query_features=torch.from_numpy(g['query_features']), upsampler_state_dict=up[1], **EXTRAS)
    lab = ours.predict(x, None)[0].cpu().numpy().astype(np.uint8)
    olg = ours.forward_slide(x, [dict(ori_shape=(H, W))], 112, 224)
    a_ours = agreement(lab, olg[0][:, ::4, ::4].cpu().numpy(), g)
    ts = []
    for _ in range(5):
        torch.cuda.synchronize()
        t0 = time.time()
        ours.predict(x, None)
        torch.cuda.synchronize()
        ts.append(time.time() - t0)
    res['ours_b200_bf16_predict'] = dict(seconds_per_tile=float(np.median(ts)), mp_per_s=H * W / 1e6 / float(np.median(ts)),
                                         vs_golden=a_ours,
                                         labels_equal_reference_fp16=float((lab == pred[0].cpu().numpy().astype(np.uint8)).mean()))
    print('[ours b200 bf16]', json.dumps(res['ours_b200_bf16_predict']))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, 'w'), indent=1)
    print('wrote', args.out)


if __name__ == '__main__':
    main()
