import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_decontamination_b200 import ops
dev = 'cuda'
M = 802816
hid = (torch.randn(M, 128, device=dev) * 0.3).bfloat16(); W = (torch.randn(128, 128, device=dev) * 0.1).bfloat16()
kern = (torch.randn(M, 128, device=dev) * 0.3).bfloat16(); out = torch.empty_like(kern); b = torch.zeros(128, device=dev)
resf = torch.randn(M, 128, device=dev)
cases = [lambda: ops.gemm(hid, W, kern, bias=b, residual=kern, alpha=0.1),      # in place (pipeline form)
         lambda: ops.gemm(hid, W, out, bias=b, residual=kern, alpha=0.1),       # separate output
         lambda: ops.gemm(hid, W, out, bias=b),                                 # no residual
         lambda: ops.gemm(hid, W, out, bias=b, residual=resf, alpha=0.1)]       # fp32 residual
for _ in range(3):
    for c in cases: c()
torch.cuda.synchronize()
for i, c in enumerate(cases):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): c()
    e.record(); torch.cuda.synchronize()
    print(i, s.elapsed_time(e) / 10, 'ms')
