#!/usr/bin/env python
"""Benchmark of the dense CLIP segmentation hot path (BASELINE.json metric: megapixels/sec segmented).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): Vaihingen-shaped 512x512 synthetic tiles, ViT-B/16 (synthetic
"random-init" weights), jbu_one upsampler (C=512, radius 5), cls_vaihingen.txt (Q=K=6), prob_thd 0.1,
bg_idx 5, slide 224/112 (16 crops per tile), base_config.py extras ON, bf16.  A step = `--tiles` tiles
through normalise-on-load -> ViT -> JBU -> logits -> accumulate/argmax -> IoU histogram (+ one all-reduce of
the [3,K] int64 histogram per step when N > 1).  Every rank processes its own tiles (weak scaling).

value : tiles already resident in HBM as uint8 (device-timed, CUDA events, max over ranks)
e2e   : the same through the call mmengine's Runner makes, ``SegmentorEx.test_step(dict(inputs=[uint8 CHW ...],
        data_samples=[...]))`` (eval.py:86-87), from pinned HOST buffers, labels + histogram read back every step
roofline / rooflines / cpu_baseline : see DESIGN.md "Measurement".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

H = W = 512
CROPS = 16
WORKLOAD = ('Vaihingen-shaped 512x512 synthetic tile, ViT-B/16 + jbu_one (C=512, r=5), Q=K=6, slide 224/112 '
            '(16 crops), base_config extras ON')
METRIC = 'megapixels/sec segmented (ViT-B/16, 512x512 tiles, jbu_one)'
# other BASELINE.json configs, selectable with --workload (the default above is the benchmark line)
WORKLOADS = {
    'vaihingen512': dict(H=512, W=512, model='ViT-B/16', cls='vaihingen', thd=0.1, bg=5, up=True),
    'potsdam512': dict(H=512, W=512, model='ViT-B/16', cls='potsdam', thd=0.1, bg=5, up=True),
    'isaid896': dict(H=896, W=896, model='ViT-B/16', cls='isaid', thd=0.4, bg=0, up=True),
    'road1024': dict(H=1024, W=1024, model='ViT-B/16', cls='roadval', thd=0.7, bg=0, up=True),
    'vaihingen512_noup': dict(H=512, W=512, model='ViT-B/16', cls='vaihingen', thd=0.1, bg=5, up=False),
    # BASELINE config 3: ViT-L/14 (L = 257, d = 1024), no upsampler (JBU x16 is shape-incompatible with patch 14), 81 crops
    'loveda1024_vitl': dict(H=1024, W=1024, model='ViT-L-14', cls='loveda', thd=0.3, bg=0, up=False,
                            text='seg_loveda_vitl_1024.npz'),
}


def _peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], tf_burst=d['bf16_tflops'], tf_sust=d.get('bf16_tflops_sustained', d['bf16_tflops']),
                    src='measured (MEASURED_PEAKS.json)')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src='fallback (B200_PROFILING.md)')


def _traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of each kernel class, from the committed
    `ncu --set full` capture of this command (profiles/r02_traffic.json; absent -> null)."""
    p = os.path.join(ROOT, 'profiles', 'r02_traffic.json')
    return json.load(open(p)) if os.path.exists(p) else {}


# ---- algorithmic work per launch of each kernel class (DESIGN.md "Kernels and their rooflines") -------
def _work_jbu_apply(n, h, w, C, radius, esize):
    """adaptive conv at the reference op boundary: padded source + kernel + output."""
    d = 2 * radius + 1
    H2, W2 = 2 * h, 2 * w
    return n * (C * (H2 + 2 * radius) * (W2 + 2 * radius) * esize + H2 * W2 * d * d * esize + C * H2 * W2 * esize)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={q}',
                                          '--format=csv,noheader,nounits', '-lms', '25'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i] == 'Active' for r in self.rows)]
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, samples=len(sm))


def build_model(device, precision='bf16', wl=None):
    from clip_decontamination_b200.open_clip import create_model
    from clip_decontamination_b200.open_clip.synthetic import synthetic_jbu_state_dict
    from clip_decontamination_b200.segmentor import SegmentorEx
    wl = wl or WORKLOADS['vaihingen512']
    if wl.get('text'):                 # query features of this model width live in the config's own golden file
        qf = np.load(os.path.join(ROOT, 'tests', 'golden', wl['text']))['query_features']
    else:
        qf = np.load(os.path.join(ROOT, 'tests', 'golden', 'bench_text.npz'))[f"{wl['cls']}_query_features"]
    net = create_model(wl['model'], pretrained=None, precision='fp32' if precision == 'fp32' else 'fp16')
    return SegmentorEx(clip_type='CLIP', vit_type=wl['model'], model_type='Experimental',
                       name_path=os.path.join(ROOT, 'configs', f"cls_{wl['cls']}.txt"), device=device,
                       prob_thd=wl['thd'], bg_idx=wl['bg'], apply_sim_feat_up=wl['up'], global_debias_factor=0.2,
                       apply_outlier_suppression=True, outlier_suppression_cfg=dict(top_k=30),
                       apply_similarity_enhancement=True,
                       similarity_enhancement_cfg=dict(similarity_weight=1.0, temperature=1.0, add_self_similarity=True),
                       sim_feat_up_cfg=dict(model_name='jbu_one', model_path=None), precision=precision, net=net,
                       query_features=torch.from_numpy(qf),
                       upsampler_state_dict=synthetic_jbu_state_dict('jbu_one', 512, 1) if wl['up'] else None)


# ---- the CPU arm: the reference's algorithm for this path on the host cores ----------------------------
EXTRAS_REF = dict(global_debias_factor=0.2, apply_outlier_suppression=True, outlier_suppression_cfg=dict(top_k=30),
                  apply_similarity_enhancement=True,
                  similarity_enhancement_cfg=dict(similarity_weight=1.0, temperature=1.0, add_self_similarity=True))


class CpuTile:
    """One FULL 512x512 tile of the workload through the CPU implementation of the path: 16 crops (ViT-B/16 + extras
    + jbu_one + cosine logits) + overlap accumulate + post-process, fp32, all host threads.  kind 'port' = the oracle
    (oracle/clipseg_oracle.py, pinned against the reference); kind 'reference' = the UNMODIFIED reference staged in
    baseline/_ref (oracle/stage_ref.py), selected with CLIPSEG_REF_ARM=unmodified -- it is ~5x slower than the port
    (tap-loop adaptive conv, per-crop Python overhead), so the port is the conservative default."""

    def __init__(self, threads):
        from clip_decontamination_b200 import synth
        from clip_decontamination_b200.open_clip.model_configs import get_model_config
        from clip_decontamination_b200.open_clip.synthetic import synthetic_clip_state_dict, synthetic_jbu_state_dict
        torch.set_num_threads(threads)
        self.threads = threads
        cfg = get_model_config('ViT-B-16')
        v = cfg['vision_cfg']
        gold = np.load(os.path.join(ROOT, 'tests', 'golden', 'bench_text.npz'))
        qf = torch.from_numpy(gold['vaihingen_query_features'])
        jbu = synthetic_jbu_state_dict('jbu_one', 512, 1)
        self.img = torch.from_numpy(synth.preprocess(synth.voronoi_scene(H, W, 100)))[None]
        self.kind = 'port'
        if os.environ.get('CLIPSEG_REF_ARM', '') == 'unmodified':
            from oracle import ref_harness as rh
            if rh.available():
                sd = synthetic_clip_state_dict(cfg, 0)
                seg = rh.build_ref_segmentor(cfg, sd, os.path.join(ROOT, 'configs', 'cls_vaihingen.txt'),
                                             model_type='Experimental', prob_thd=0.1, bg_idx=5,
                                             upsampler=('jbu_one', jbu), **EXTRAS_REF)
                seg.query_features = qf

                def run():
                    lg = seg.forward_slide(self.img, [dict(ori_shape=(H, W))], 112, 224)
                    return seg.postprocess_result(lg, None)
                self.run, self.kind = run, 'reference'
                return
        from oracle import clipseg_oracle as O
        sd = synthetic_clip_state_dict(cfg, 0, text_tower=False)
        vis = {k[len('visual.'):]: t for k, t in sd.items() if k.startswith('visual.')}
        orc = O.SegOracle(vis, qf, list(range(6)), layers=v['layers'], heads=v['heads'], patch=16, prob_thd=0.1, bg_idx=5,
                          global_debias_factor=0.2, upsampler=('jbu_one', jbu), sim_cfg={}, outlier_cfg={'top_k': 30})
        self.run = lambda: orc.predict(self.img)

    def seconds(self):
        with torch.no_grad():
            t0 = time.time()
            self.run()
            return time.time() - t0

    @property
    def sample(self):
        return ('one full 512x512 tile per step: 16 crops (ViT-B/16 + extras + jbu_one + cosine logits) + overlap '
                'accumulate + post-process, fp32, ' + ('unmodified reference (baseline/_ref)' if self.kind == 'reference'
                                                       else 'oracle port of the reference'))


def run_reference(args, rank):
    """--impl reference: the CPU implementation of the path on the host cores, all host threads; a step = one full
    tile (no extrapolation); exactly --steps timed steps after --warmup warm-up steps."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    cpu = CpuTile(threads)
    for _ in range(max(0, args.warmup)):
        cpu.seconds()
    ts = [cpu.seconds() for _ in range(max(1, args.steps))]
    sec_tile = float(np.mean(ts))
    mps = H * W / 1e6 / sec_tile
    line = dict(metric=METRIC, value=mps, unit='MP/s', n_gpus=args.gpus, steps=len(ts), warmup=max(0, args.warmup),
                ms_per_step=sec_tile * 1e3, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
                data='synthetic', impl='reference', config=dict(workload=WORKLOAD, tiles_per_step=1),
                cpu_baseline=dict(value=mps, unit='MP/s', cores=threads, kind=cpu.kind, sample=cpu.sample),
                e2e=dict(value=mps, unit='MP/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--tiles', type=int, default=6, help='tiles per step per GPU')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--workload', default='vaihingen512', choices=sorted(WORKLOADS))
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if args.impl == 'reference':
        return run_reference(args, rank)
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    from clip_decontamination_b200 import ops, synth
    from clip_decontamination_b200 import _lib
    from clip_decontamination_b200.compat import SegDataSample
    from clip_decontamination_b200.dist import allreduce_hist
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=device)
    global H, W, CROPS, WORKLOAD, METRIC
    wl = WORKLOADS[args.workload]
    if args.workload != 'vaihingen512':
        from clip_decontamination_b200.engine import slide_windows
        H, W = wl['H'], wl['W']
        CROPS = len(slide_windows(H, W, 112, 224))
        WORKLOAD = f"{args.workload}: {H}x{W} synthetic tile, {wl['model']}, cls_{wl['cls']}.txt, jbu_one={wl['up']}, {CROPS} crops, extras ON"
        args.no_cpu_baseline = True
        METRIC = f"megapixels/sec segmented ({wl['model']}, {H}x{W} tiles, {'jbu_one' if wl['up'] else 'no upsampler'})"
    model = build_model(device, args.precision, wl)
    eng = model.engine
    K = model.num_classes
    T = args.tiles
    # synthetic tiles (different per rank / tile) as the dataloader hands them over: uint8 CHW BGR (PackSegInputs),
    # one pinned host tensor per image; synthetic ground truth for the histogram
    host_imgs = [torch.from_numpy(np.ascontiguousarray(synth.voronoi_scene(H, W, 1000 + rank * 64 + t).transpose(2, 0, 1))).pin_memory()
                 for t in range(T)]
    host_gt = torch.stack([torch.from_numpy(synth.synthetic_labels(H, W, K, 2000 + rank * 64 + t)) for t in range(T)])
    dev_imgs = torch.stack(host_imgs).to(device)        # [T,3,H,W] uint8 BGR
    dev_gt = host_gt.to(device)
    hist = torch.zeros((3, K), dtype=torch.int64, device=device)
    labels = torch.empty((T, H, W), dtype=torch.uint8, device=device)
    # e2e results land in double-buffered pinned host memory; the result of step i is awaited while step i+1 is in flight
    host_labels = [torch.empty((T, H, W), dtype=torch.uint8).pin_memory() for _ in range(2)]
    host_hist = [torch.empty((3, K), dtype=torch.int64).pin_memory() for _ in range(2)]
    d2h_done = [torch.cuda.Event() for _ in range(2)]
    e2e_state = dict(i=0, pending=None, checksum=0)
    samples = [SegDataSample(dict(ori_shape=(H, W), img_shape=(H, W))) for _ in range(T)]
    dev_image = ops.Image.u8(dev_imgs, 'chw', eng.mean, eng.std)

    # One step = one pass of the hot path over one batch of T tiles: the T images go through every kernel together
    # (T x 16 crops per launch), as the reference's slide_inference does with a batched input.
    def step_eager():          # un-graphed launch sequence (used for the instrumented breakdown)
        eng.segment(dev_image, None, labels=labels.view(T * H, W))
        ops.iou_hist(labels.view(-1), dev_gt.view(-1), K, hist)
        allreduce_hist(hist)

    def step_resident():       # the product path: CUDA-graph replay, inputs resident in HBM
        lab = eng.segment_batch(dev_imgs, 'u8chw', copy_out=False)
        ops.iou_hist(lab.view(-1), dev_gt.view(-1), K, hist)
        allreduce_hist(hist)

    def e2e_consume(slot):     # the host reads the finished result of an earlier step (pinned memory, after its D2H event)
        d2h_done[slot].synchronize()
        e2e_state['checksum'] += int(host_hist[slot][1].sum()) + int(host_labels[slot][0, 0, 0])

    def step_e2e():            # what mmengine's Runner.test() does per batch (eval.py:86-87) + the metric
        slot = e2e_state['i'] & 1
        out = model.test_step(dict(inputs=host_imgs, data_samples=samples))      # H2D of the raw bytes inside
        lab = model.last_labels                                                    # uint8 [T,H,W] behind pred_sem_seg
        ops.iou_hist(lab.view(-1), dev_gt.view(-1), K, hist)
        allreduce_hist(hist)
        host_labels[slot].copy_(lab, non_blocking=True)                            # D2H of the step's result
        host_hist[slot].copy_(hist, non_blocking=True)
        d2h_done[slot].record()
        if e2e_state['pending'] is not None:                                       # read step i-1 while step i runs
            e2e_consume(e2e_state['pending'])
        e2e_state['pending'] = slot
        e2e_state['i'] += 1
        return out

    def e2e_drain():
        if e2e_state['pending'] is not None:
            e2e_consume(e2e_state['pending'])
            e2e_state['pending'] = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, drain=None):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        if drain is not None:
            drain()                                   # every step's result has been read by the host inside the region
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- warm-up; find the kernel classes of a step with a fully instrumented eager step ----------------
    for _ in range(args.warmup):
        step_eager()
        step_resident()
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    step_eager()
    launches_per_step = _lib.launch_count() - l0       # kernels per step (a graph replay issues the same set)
    records = {}
    shape_records = {}
    orig = {}

    # kernel classes: the shared-kernel entry points belong to the class of the op they implement
    CLASS = dict(jbu_apply_shared='jbu_apply', jbu_composite_image='jbu_apply', jbu_range_kernel_border='jbu_range_kernel',
                 attention_experimental_tc='attention', gemm_blockdiag='gemm')

    def wrap(name, fn, workfn=None):
        cname = CLASS.get(name, name)

        def inner(*a, **k):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = fn(*a, **k)
            e.record()
            records.setdefault(cname, []).append((s, e, workfn(*a, **k) if workfn else None))
            if name == 'gemm':      # per-shape split of the GEMM class (diagnostic)
                A, B = a[0], a[1]
                key = 'gemm[M%dxN%dxK%d]' % (k.get('M') or A.shape[0], k.get('N') or B.shape[0], k.get('K') or A.shape[1])
                Mg, Ng, Kg = k.get('M') or A.shape[0], k.get('N') or B.shape[0], k.get('K') or A.shape[1]
                # bytes at the op boundary: both operands once, the output, the residual (fp32 for the residual stream)
                nbytes = Mg * Kg * A.element_size() + Ng * Kg * B.element_size()
                if len(a) > 2 and torch.is_tensor(a[2]):
                    nbytes += Mg * Ng * a[2].element_size()
                if torch.is_tensor(k.get('residual')):
                    nbytes += Mg * Ng * k['residual'].element_size()
                shape_records.setdefault(key, []).append((s, e, 2.0 * Mg * Ng * Kg, float(nbytes)))
            return r
        return inner

    esize = 2 if args.precision == 'bf16' else 4

    def gemm_work(A, B, out, **k):
        return ('tensor', 2.0 * (k.get('M') or A.shape[0]) * (k.get('N') or B.shape[0]) * (k.get('K') or A.shape[1]))

    def apply_work(src, n, h, w, C, kern, radius, dst, hr, *a, **k):
        return ('hbm', float(_work_jbu_apply(n, h, w, C, radius, esize)))

    def apply_shared_work(src, n, h, w, C, kern_b, kern_img, kc_img, windows, shift, pitch, radius, dst, scratch):
        return ('hbm', float(_work_jbu_apply(n, h, w, C, radius, esize)))      # same op boundary, all n crops

    def comp_img_work(kern_img, ih, iw, gh, gw, radius, kc_img, tabs):
        return ('hbm', 0.0)                         # part of the apply class: its time counts, the op-boundary bytes do not change

    def rkb_work(proj, guid, windows, shift, pitch, n, gh, gw, radius, rt, ss, kern_b):
        return ('hbm', float(kern_b.shape[0] * (32 * proj.element_size() + 16 + kern_b.shape[-1] * esize)))   # rows computed

    def attn_work(qkv, n, L, heads, hd, mode, out, **k):
        return ('tensor', (4.0 if mode == 0 else 6.0) * n * heads * L * L * hd)

    def gemm_bd_work(A, B, out, block_rows):
        return ('tensor', 2.0 * A.shape[0] * block_rows * A.shape[1])       # the diagonal blocks only

    def attn_exp_work(qkv, n, L, heads, out, *a, **k):
        return ('tensor', 6.0 * n * heads * L * L * 64)

    def nsim_work(feats, ldf, n, hw, D, text, logits, cls_logit_bias=None):
        return ('hbm', float(n * hw * (D * esize + text.shape[0] * 4)))

    def accum_work(cl, *a, **k):
        return ('hbm', float(cl.numel() * 4 + T * H * W))

    def rk_work(proj, guid, n, gh, gw, radius, rt, ss, kern, *a, **k):
        rows = k.get('n_rows') or n * gh * gw
        return ('hbm', float(rows * (32 * proj.element_size() + 16 + kern.shape[-1] * esize)))

    def fns_work(y, Wt, n, hw, Cc, bias, alpha, text, logits, cls_logit_bias=None, scratch=None):
        return ('tensor', 2.0 * n * hw * Cc * (Cc + text.shape[0]))

    def basis_work(s, Cb, n, hw, Tk, ts, gram, aux, consts, Q, logits, cls_logit_bias=None):
        return ('hbm', float(n * hw * (Cb * esize + Q * 4)))

    def kfix_work(k, W0, b0, W3s, b3s, out):
        return ('hbm', float(2 * k.shape[0] * k.shape[1] * esize))

    def ln_work(x, gamma, beta, out, eps=1e-5):
        return ('hbm', float(x.numel() * (4 + out.element_size())))

    def simmap_work(x, n, L, width, out, *a, **k):
        return ('hbm', float(n * L * width * 4 + n * (L - 1) * (L - 1) * 4))

    workfns = dict(jbu_kernel_fixup=kfix_work, basis_logits=basis_work, fixup_norm_sim=fns_work, gemm=gemm_work,
                   jbu_apply=apply_work, attention=attn_work, norm_sim=nsim_work, accum_argmax=accum_work,
                   jbu_range_kernel=rk_work, layernorm=ln_work, simmap=simmap_work, jbu_apply_shared=apply_shared_work,
                   jbu_composite_image=comp_img_work, jbu_range_kernel_border=rkb_work,
                   attention_experimental_tc=attn_exp_work, gemm_blockdiag=gemm_bd_work)
    names = ['preprocess_u8', 'patchify', 'embed_tokens', 'embed_tokens_ln', 'layernorm', 'gemm', 'gemm_blockdiag', 'attention',
             'attention_experimental_tc', 'simmap', 'outlier_suppress',
             'cls_debias', 'jbu_guidance', 'jbu_range_proj', 'jbu_guidance_proj', 'jbu_range_kernel', 'jbu_kernel_fixup',
             'jbu_apply', 'jbu_apply_shared', 'jbu_composite_image', 'jbu_range_kernel_border', 'norm_sim', 'fixup_norm_sim',
             'basis_logits', 'accum_argmax', 'iou_hist']
    names = [nm for nm in names if hasattr(ops, nm)]
    for nm in names:
        orig[nm] = getattr(ops, nm)
        setattr(ops, nm, wrap(nm, orig[nm], workfns.get(nm)))
    # the instrumented passes run every kernel ALONE on the stream (no side-stream branch next to the ViT): a launch
    # bracketed by events while another stream shares the SMs would be charged for its neighbour's work
    overlap = eng.overlap_image_level
    eng.overlap_image_level = False
    step_eager()
    torch.cuda.synchronize()
    for nm in names:
        setattr(ops, nm, orig[nm])
    totals = {nm: sum(s.elapsed_time(e) for s, e, _ in recs) for nm, recs in records.items()}
    tot_all = sum(totals.values())
    breakdown = {nm: round(v / T, 4) for nm, v in sorted(totals.items(), key=lambda kv: -kv[1])}
    shape_records.clear()
    classes = [nm for nm in totals if nm in workfns and totals[nm] >= 0.03 * tot_all]     # every class >= 3 % of a step
    members = {c: [nm for nm in names if CLASS.get(nm, nm) == c] for c in classes}
    dominant = max(classes, key=lambda nm: totals[nm])
    records.clear()
    # ---- timed region: value (inputs resident in HBM, CUDA-graph replay of the launch sequence) -----
    clk = ClockSampler(local_rank)
    clk.start()
    eng.overlap_image_level = overlap
    ms = timed(step_resident, args.steps)
    launches = launches_per_step * args.steps          # a replay issues the kernels counted at capture
    eng.overlap_image_level = False
    # ---- the same launches issued eagerly with every kernel class >= 3 % of the step bracketed by CUDA events (a graph
    #      node cannot be bracketed): per-launch durations for the rooflines -------------------------------------------
    for c in classes:
        for nm in members[c]:
            setattr(ops, nm, wrap(nm, orig[nm], workfns[nm]))
    ms_eager = timed(step_eager, args.steps)
    eng.overlap_image_level = overlap
    clocks = clk.stop()
    for c in classes:
        for nm in members[c]:
            setattr(ops, nm, orig[nm])
    peaks = _peaks()
    traffic = _traffic()

    def roofline_of(nm):
        recs = records.get(nm, [])
        t_ms = float(sum(s.elapsed_time(e) for s, e, _ in recs))
        work = float(sum(w[1] for _, _, w in recs))
        bound = recs[0][2][0] if recs else 'hbm'
        if bound == 'hbm':
            achieved, peak, runit = work / t_ms / 1e6, peaks['hbm'], 'GB/s'
        else:
            achieved, peak, runit = work / t_ms / 1e9, peaks['tf_sust'], 'TFLOP/s'
        return dict(kernel=nm, bound=bound, achieved=achieved, peak=peak, unit=runit, frac=achieved / peak,
                    traffic=traffic.get(nm), launches_timed=len(recs), avg_launch_ms=t_ms / max(1, len(recs)),
                    share_of_step=t_ms / ms_eager)

    rooflines = sorted((roofline_of(nm) for nm in classes), key=lambda r: -r['share_of_step'])
    def shape_row(v):
        # a GEMM shape is bound by whichever roofline it sits closer to: out-proj (K = width, fp32 residual in and out) moves
        # 145 MB for 22 GFLOP at the bench batch and is HBM bound, QKV / fc1 / fc2 are tensor bound
        t_ms = sum(s.elapsed_time(e) for s, e, _, _ in v)
        tf, gbs = sum(w for _, _, w, _ in v) / t_ms / 1e9, sum(b for _, _, _, b in v) / t_ms / 1e6
        ft, fh = tf / peaks['tf_sust'], gbs / peaks['hbm']
        return dict(calls_per_step=len(v) // args.steps, ms_per_step=round(t_ms / args.steps, 4), tflops=round(tf, 1),
                    gbs_at_op_boundary=round(gbs, 1), bound='tensor' if ft >= fh else 'hbm', frac=round(max(ft, fh), 3))

    gemm_shapes = {k: shape_row(v) for k, v in shape_records.items()}
    roofline = dict(next(r for r in rooflines if r['kernel'] == dominant))
    roofline.update(eager_ms_per_step=ms_eager / args.steps, peak_source=peaks['src'], per_tile_ms_by_kernel=breakdown,
                    gemm_by_shape=gemm_shapes)

    # ---- e2e: host buffers in, labels + histogram out, through SegmentorEx.test_step ----------------
    for _ in range(2):
        step_e2e()
    e2e_drain()
    ms_e2e = timed(step_e2e, args.steps, e2e_drain)

    mp_step = T * H * W / 1e6 * world
    value = mp_step * args.steps / (ms / 1e3)
    e2e = mp_step * args.steps / (ms_e2e / 1e3)
    line = dict(metric=METRIC, value=value, unit='MP/s', n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype=args.precision, data='synthetic',
                config=dict(workload=WORKLOAD, tiles_per_step_per_gpu=T, crops_per_tile=CROPS,
                            l2='working set of a step (JBU stage buffers, ~95 MB per crop: ~9 GB for the 96 crops of a step) far exceeds the 126 MB L2; '
                               'the batch holds %d different tiles' % T, parallelism=f'image-sharded x{world}'),
                clocks=clocks, gpu_launches=int(launches),
                e2e=dict(value=e2e, unit='MP/s', h2d_bytes_per_step=T * H * W * 3, d2h_bytes_per_step=T * H * W + 3 * K * 8,
                         ms_per_step=ms_e2e / args.steps, api='SegmentorEx.test_step(dict(inputs=[uint8 CHW], data_samples))',
                         pipeline='H2D of step i+1 on a copy stream overlaps step i; the host reads result i-1 while step i runs'),
                roofline=roofline, rooflines=rooflines)
    ref_gpu = os.path.join(ROOT, 'profiles', 'r02_reference_on_b200.json')
    if os.path.exists(ref_gpu):       # measured once with oracle/ref_on_gpu.py on the same kind of box (stated baseline)
        rg = json.load(open(ref_gpu))
        line['reference_on_b200'] = {k: dict(mp_per_s=v['mp_per_s'], seconds_per_tile=v['seconds_per_tile'])
                                     for k, v in rg.items() if isinstance(v, dict) and 'mp_per_s' in v}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cpu = CpuTile(threads)
        sec = cpu.seconds()
        line['cpu_baseline'] = dict(value=H * W / 1e6 / sec, unit='MP/s', cores=threads, kind=cpu.kind, sample=cpu.sample)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
