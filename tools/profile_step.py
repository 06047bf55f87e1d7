"""One eager step of the bench workload between cudaProfilerStart/Stop, for Nsight Compute:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv \
      python tools/profile_step.py --tiles 6
  ncu --profile-from-start off --set full --clock-control none --import-source on -o r02_full python tools/profile_step.py --tiles 1

(after the same command has exited 0 without ncu).  A step = the launch sequence bench.py replays from its CUDA graph."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from clip_decontamination_b200 import ops, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--tiles', type=int, default=6)
ap.add_argument('--workload', default='vaihingen512')
ap.add_argument('--layers', type=int, default=0, help='truncate the ViT to this many blocks (0 = all): one launch per GEMM shape for --set full')
args = ap.parse_args()
wl = bench.WORKLOADS[args.workload]
H, W, T = wl['H'], wl['W'], args.tiles
torch.cuda.set_device(0)
model = bench.build_model(torch.device('cuda', 0), 'bf16', wl)
eng, K = model.engine, model.num_classes
if args.layers:                       # first blocks + the final block (the modified attention lives in the last one)
    v = eng.v
    v.blocks = v.blocks[:args.layers - 1] + v.blocks[-1:]
    v.layers = args.layers
imgs = torch.stack([torch.from_numpy(np.ascontiguousarray(synth.voronoi_scene(H, W, 1000 + t).transpose(2, 0, 1))) for t in range(T)]).cuda()
gt = torch.stack([torch.from_numpy(synth.synthetic_labels(H, W, K, 2000 + t)) for t in range(T)]).cuda()
hist = torch.zeros((3, K), dtype=torch.int64, device='cuda')
labels = torch.empty((T, H, W), dtype=torch.uint8, device='cuda')
image = ops.Image.u8(imgs, 'chw', eng.mean, eng.std)


def step():
    eng.segment(image, None, labels=labels.view(T * H, W))
    ops.iou_hist(labels.view(-1), gt.view(-1), K, hist)


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('profiled one step of', T, 'tiles', H, 'x', W)
