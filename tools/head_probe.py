"""Final-stage head kernels in isolation (for ncu): python tools/head_probe.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_decontamination_b200 import ops
n, T, Cb, Q, hw, ldk = 16, 196, 256, 6, 224 * 224, 128
dev = 'cuda'
Tp = 200
S = torch.rand(n * hw, Cb, device=dev).bfloat16()
ldg = n * Tp + 8
gram = torch.randn(n * Tp, ldg, device=dev).bfloat16()
aux = torch.randn(16, ldg, device=dev).bfloat16()
consts = torch.randn(Q + 1, device=dev).abs() + 1
lg = torch.empty(n, Q, hw, device=dev)
k = torch.rand(n * hw, ldk, device=dev).bfloat16()
out = torch.empty_like(k)
W0 = (torch.randn(ldk, ldk, device=dev) * 0.1).bfloat16()
W3 = (torch.randn(ldk, ldk, device=dev) * 0.01).bfloat16()
b = torch.zeros(ldk, device=dev)
for _ in range(3):
    ops.basis_logits(S, Cb, n, hw, T, Tp, gram, aux, consts, Q, lg)
    ops.jbu_kernel_fixup(k, W0, b, W3, b, out)
torch.cuda.synchronize()
print('ok')
