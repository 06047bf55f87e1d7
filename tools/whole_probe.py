"""Whole-image inference (slide_crop = 0, segmentor.py:470-471) at sizes far beyond one crop: runs the full path on one
H x W tile as a single 'crop' of L = H W / 256 + 1 tokens (long-sequence attention kernel), checks the outputs are finite
and the bf16 labels agree with the fp32 verification mode, and times it.  python tools/whole_probe.py H W [jbu]"""
import os, sys, time, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from clip_decontamination_b200 import synth
H, W = int(sys.argv[1]), int(sys.argv[2])
up = len(sys.argv) > 3 and sys.argv[3] == 'jbu'
wl = dict(bench.WORKLOADS['vaihingen512'], up=up)
img = torch.from_numpy(synth.preprocess(synth.voronoi_scene(H, W, 7))).cuda()
res = {}
for prec in ('bf16', 'fp32'):
    model = bench.build_model(torch.device('cuda', 0), prec, wl)
    eng = model.engine
    eng.crop = 0
    for it in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        labels, probs, avg = eng.segment(img, None, want_logits=True, want_probs=True)
        torch.cuda.synchronize(); dt = time.time() - t0
    assert torch.isfinite(avg).all() and torch.isfinite(probs).all()
    res[prec] = (labels.cpu(), avg.cpu())
    print(f'{H}x{W} whole image ({"jbu_one" if up else "no upsampler"}), L={H * W // 256 + 1}, {prec}: {dt * 1e3:.1f} ms, '
          f'label hist {np.bincount(labels.cpu().numpy().ravel(), minlength=eng.K).tolist()}')
    del model, eng
    torch.cuda.empty_cache()
d = (res['bf16'][1] - res['fp32'][1]).abs().max().item()
agree = (res['bf16'][0] == res['fp32'][0]).float().mean().item()
print(f'bf16 vs fp32 mode: max|dlogit| = {d:.3e}, label agreement {agree * 100:.3f}%')
assert d < 1e-2
