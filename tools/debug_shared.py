"""Step-by-step check of the shared JBU kernel generation against the per-crop form (run with
CUDA_LAUNCH_BLOCKING=1, optionally under compute-sanitizer)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from clip_decontamination_b200 import ops, synth  # noqa: E402
from clip_decontamination_b200.engine import JBUEngine, slide_windows  # noqa: E402
from clip_decontamination_b200.open_clip.synthetic import synthetic_jbu_state_dict  # noqa: E402

torch.cuda.set_device(0)
H = W = int(os.environ.get('DBG_SIZE', 448))
C = 256
eng = JBUEngine('jbu_one', synthetic_jbu_state_dict('jbu_one', 512, 1), 512, 'bf16')
img = torch.from_numpy(synth.preprocess(synth.voronoi_scene(H, W, 3))).cuda()
wl = slide_windows(H, W, 112, 224)
wins = torch.tensor(wl, dtype=torch.int32).cuda()
n = len(wl)
print('windows', n, 'share_ok', eng.share_ok(wl, H * W, 224, 224, 0, 0))


def sync(tag):
    torch.cuda.synchronize()
    print('ok', tag, flush=True)


def coords(GH, GW, fb):
    ys, xs = [], []
    for y in range(GH):
        for x in range(GW):
            if not (fb <= y < GH - fb and 16 <= x < GW - 16):
                ys.append(y); xs.append(x)
    idx = []
    for y, x in zip(ys, xs):
        if y < fb:
            idx.append(y * GW + x)
        elif y >= GH - fb:
            idx.append(fb * GW + (y - (GH - fb)) * GW + x)
        else:
            idx.append(2 * fb * GW + (y - fb) * 32 + (x if x < 16 else 16 + x - (GW - 16)))
    return torch.tensor(ys).cuda(), torch.tensor(xs).cuda(), torch.tensor(idx).cuda()


sh = eng.prepare_shared(img, 224, 224, 14, 14)
sync('prepare_shared')
for si in (2, 3):
    st = eng.stages[si]
    GH = GW = 14 << (si + 1)
    h = w = GH // 2
    npix = n * GH * GW
    kw = st['ldk']
    # per-crop reference
    guid = torch.empty((npix, 4), device='cuda')
    proj = torch.empty((npix, 32), dtype=torch.float16, device='cuda')
    ops.jbu_guidance_proj(img, wins, 224, 224, 0, 0, GH, GW, st['rp_w0'], st['rp_b0'], st['rp_w3'], st['rp_b3'], guid, proj)
    kraw = torch.empty((npix, kw), dtype=torch.bfloat16, device='cuda')
    kern = torch.empty_like(kraw)
    ops.jbu_range_kernel(proj, guid, n, GH, GW, st['radius'], st['range_temp'], st['sigma'], kraw)
    ops.jbu_kernel_fixup(kraw, st['fx_w0'], st['fx_b0'], st['fx_w3'][:, :kw], st['fx_b3'], kern)
    src = torch.randn((n * h * w, C), device='cuda').to(torch.bfloat16)
    hr = torch.zeros((npix, C), dtype=torch.bfloat16, device='cuda')
    dst = torch.empty((npix, C), dtype=torch.bfloat16, device='cuda')
    ops.jbu_apply(src, n, h, w, C, kern, st['radius'], dst, hr)
    sync(f'stage {si}: per-crop')
    kc_ref = hr.view(-1)[:npix * 128].view(n, GH, GW, 128).clone()
    s = sh[si]
    shift, pitch = s['shift'], s['pitch']
    IH, IW = H >> shift, W >> shift
    # image-level vs per-crop on interior pixels
    gi = s['guid'].view(IH, IW, 4)
    pi = s['proj'].view(IH, IW, 32)
    ki = s['kern'].view(IH, IW, kw)
    kci = s['kc'].view(IH, IW, 128)
    for c, (y1, x1, _, _) in enumerate(wl):
        oy, ox = y1 >> shift, x1 >> shift
        g_c = guid.view(n, GH, GW, 4)[c]
        assert torch.equal(gi[oy:oy + GH, ox:ox + GW], g_c), 'guidance differs'
        assert torch.equal(pi[oy:oy + GH, ox:ox + GW], proj.view(n, GH, GW, 32)[c]), 'proj differs'
        k_c = kern.view(n, GH, GW, kw)[c]
        d = (ki[oy + 8:oy + GH - 8, ox + 16:ox + GW - 16].float() - k_c[8:GH - 8, 16:GW - 16].float()).abs().max().item()
        dk = (kci[oy + 12:oy + GH - 12, ox + 16:ox + GW - 16, :81].float() - kc_ref[c, 12:GH - 12, 16:GW - 16, :81].float()).abs().max().item()
        if c < 3 or d > 0 or dk > 0:
            print(f'  stage {si} crop {c}: interior kern max|d| = {d:.3e}  composite max|d| = {dk:.3e}')
    # border tensors
    rb = ops.jbu_share_rows(GH, GW, ops.FB_RANGE)
    kraw_b = torch.zeros((n * rb, kw), dtype=torch.bfloat16, device='cuda')
    kern_b = torch.zeros_like(kraw_b)
    ops.jbu_range_kernel_border(s['proj'], s['guid'], wins, shift, pitch, n, GH, GW, st['radius'], st['range_temp'], st['sigma'], kraw_b)
    sync(f'stage {si}: range border')
    ys, xs, idx = coords(GH, GW, ops.FB_RANGE)
    for c in range(n):
        d = (kraw_b.view(n, rb, kw)[c, idx].float() - kraw.view(n, GH, GW, kw)[c, ys, xs].float()).abs().max().item()
        if c < 3 or d > 0:
            print(f'  stage {si} crop {c}: border kraw max|d| = {d:.3e}')
    ops.jbu_kernel_fixup(kraw_b, st['fx_w0'], st['fx_b0'], st['fx_w3'][:, :kw], st['fx_b3'], kern_b)
    sync(f'stage {si}: fixup border')
    hr2 = torch.zeros((npix, C), dtype=torch.bfloat16, device='cuda')
    dst2 = torch.empty((npix, C), dtype=torch.bfloat16, device='cuda')
    ops.jbu_apply_shared(src, n, h, w, C, kern_b, s['kern'], s['kc'], wins, shift, pitch, st['radius'], dst2, hr2)
    sync(f'stage {si}: apply shared')
    rbc = ops.jbu_share_rows(GH, GW, ops.FB_COMP)
    kcb = hr2.view(-1)[:n * rbc * 128].view(n, rbc, 128)
    ys, xs, idx = coords(GH, GW, ops.FB_COMP)
    for c in range(n):
        d = (kcb[c, idx, :81].float() - kc_ref[c, ys, xs, :81].float()).abs().max().item()
        if c < 3 or d > 0:
            print(f'  stage {si} crop {c}: border composite max|d| = {d:.3e}')
    print(f'stage {si}: dst max|d| = {(dst.float() - dst2.float()).abs().max().item():.3e}', flush=True)
print('done')
