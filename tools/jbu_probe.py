"""JBU last-stage kernels in isolation (for ncu): python tools/jbu_probe.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_decontamination_b200 import ops
n, C = 16, 512
dev = 'cuda'
src = torch.randn(n * 112 * 112, C, device=dev).bfloat16()
kern = torch.rand(n * 224 * 224, 128, device=dev).bfloat16()
dst = torch.empty(n * 224 * 224, C, device=dev, dtype=torch.bfloat16)
hr = torch.empty_like(dst)
proj = torch.randn(n * 224 * 224, 32, device=dev).half()
guid = torch.randn(n * 224 * 224, 4, device=dev)
text = torch.randn(6, C, device=dev)
lg = torch.empty(n, 6, 224 * 224, device=dev)
for _ in range(3):
    ops.jbu_apply(src, n, 112, 112, C, kern, 5, dst, hr)
    ops.jbu_range_kernel(proj, guid, n, 224, 224, 5, 0.3, 1.0, kern)
    ops.norm_sim(dst, C, n, 224 * 224, C, text, lg)
torch.cuda.synchronize()
print('ok')
