"""GEMM shapes of the pipeline in isolation, with their real epilogues (for ncu): python tools/gemm_probe.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_decontamination_b200 import ops
dev = 'cuda'
def mk(M, N, K):
    return (torch.randn(M, K, device=dev) * 0.3).bfloat16(), (torch.randn(N, K, device=dev) * 0.1).bfloat16()
cases = []
A, B = mk(802816, 128, 128); C = torch.empty(802816, 128, device=dev, dtype=torch.bfloat16); b = torch.zeros(128, device=dev)
cases.append(lambda: ops.gemm(A, B, C, bias=b, act=1))                          # jbu fix-up 1 (GELU)
cases.append(lambda: ops.gemm(C, B, A, bias=b, residual=A, alpha=0.1))           # jbu fix-up 2 (in-place residual)
A2, B2 = mk(3152, 3072, 768); C2 = torch.empty(3152, 3072, device=dev, dtype=torch.bfloat16); b2 = torch.zeros(3072, device=dev)
cases.append(lambda: ops.gemm(A2, B2, C2, bias=b2, act=1))                       # fc1
B3 = (torch.randn(768, 3072, device=dev) * 0.1).bfloat16(); x = torch.randn(3152, 768, device=dev); b3 = torch.zeros(768, device=dev)
cases.append(lambda: ops.gemm(C2, B3, x, bias=b3, residual=x))                   # fc2 (fp32 residual in place)
A4, B4 = mk(802816, 512, 512); C4 = torch.empty(802816, 512, device=dev, dtype=torch.bfloat16); b4 = torch.zeros(512, device=dev)
cases.append(lambda: ops.gemm(A4, B4, C4, bias=b4, residual=A4, alpha=0.1))      # final 1x1 conv
for _ in range(3):
    for c in cases: c()
torch.cuda.synchronize()
for i, c in enumerate(cases):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): c()
    e.record(); torch.cuda.synchronize()
    print(i, s.elapsed_time(e) / 10, 'ms')
