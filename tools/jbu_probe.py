"""JBU last-stage kernels in isolation (for ncu): python tools/jbu_probe.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_decontamination_b200 import ops
n, C = 16, int(os.environ.get('PROBE_C', 256))   # 256 = basis width of the bf16 pipeline (196 tokens)
dev = 'cuda'
src = torch.randn(n * 112 * 112, C, device=dev).bfloat16()
kern = torch.rand(n * 224 * 224, 128, device=dev).bfloat16()
dst = torch.empty(n * 224 * 224, C, device=dev, dtype=torch.bfloat16)
hr = torch.empty_like(dst)
proj = torch.randn(n * 224 * 224, 32, device=dev).half()
guid = torch.randn(n * 224 * 224, 4, device=dev)
for _ in range(3):
    ops.jbu_apply(src, n, 112, 112, C, kern, 5, dst, hr)
    ops.jbu_range_kernel(proj, guid, n, 224, 224, 5, 0.3, 1.0, kern)
torch.cuda.synchronize()
print('ok')
