import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_decontamination_b200 import ops
dev = 'cuda'
A = (torch.randn(3152, 3072, device=dev) * 0.3).bfloat16(); B = (torch.randn(768, 3072, device=dev) * 0.1).bfloat16()
C = torch.empty(3152, 768, device=dev)
A2 = (torch.randn(8192, 8192, device=dev) * 0.3).bfloat16(); B2 = (torch.randn(8192, 8192, device=dev) * 0.1).bfloat16()
C2 = torch.empty(8192, 8192, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    ops.gemm(A, B, C)
    ops.gemm(A2, B2, C2)
torch.cuda.synchronize()
print('ok')
