import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_decontamination_b200 import ops
from clip_decontamination_b200._lib import ATTN
n, L, heads, hd = 16, 197, 12, 64
qkv = torch.randn(n * L, 3 * heads * hd, device='cuda').bfloat16()
out = torch.empty(n * L, heads * hd, device='cuda', dtype=torch.bfloat16)
for _ in range(5):
    ops.attention(qkv, n, L, heads, hd, ATTN['STD'], out)
torch.cuda.synchronize()
print('ok')
