"""Timing probe of the attention kernels (CUDA events, 20 launches each): python tools/attn_probe.py n L heads"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_decontamination_b200 import ops
from clip_decontamination_b200._lib import ATTN
n, L, heads = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (96, 197, 12)
hd = 64
qkv = torch.randn(n * L, 3 * heads * hd, device='cuda').bfloat16()
out = torch.empty(n * L, heads * hd, device='cuda', dtype=torch.bfloat16)
stats = torch.zeros((n, heads, 2, L - 1), device='cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def t(fn, it=20):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(it):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / it * 1e3


flops = 4.0 * L * L * hd * heads * n
for name, fn in [('STD', lambda: ops.attention(qkv, n, L, heads, hd, ATTN['STD'], out)),
                 ('STD+stats', lambda: ops.attention(qkv, n, L, heads, hd, ATTN['STD'], out, stats=stats)),
                 ('vanilla (mma.sync)', lambda: ops.attention(qkv, n, L, heads, hd, ATTN['vanilla'], out))]:
    us = t(fn)
    print(f'n={n} L={L} heads={heads} {name}: {us:.1f} us  {flops / us * 1e-6:.1f} TFLOP/s  '
          f'{(n * L * heads * hd * 2 * 4) / us * 1e-3:.0f} GB/s at the op boundary')
if L <= ops.SIMT_COLS_MAX:
    simt = torch.rand(n, ops.simt_floats(L), device='cuda')
    us = t(lambda: ops.attention_experimental_tc(qkv, n, L, heads, out, simt, 1.0))
    print(f'n={n} L={L} heads={heads} Experimental (tcgen05, similarity map): {us:.1f} us  {1.5 * flops / us * 1e-6:.1f} TFLOP/s')
    us = t(lambda: ops.attention_experimental_tc(qkv, n, L, heads, out, None, 1.0))
    print(f'n={n} L={L} heads={heads} Experimental (tcgen05, no map): {us:.1f} us')
