"""Micro-timings (CUDA events) of individual libclipseg kernels on the shapes of BASELINE config 1.
Not the benchmark (bench.py is); used to decide what to optimise next.  python tools/kernel_probe.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_decontamination_b200 import ops  # noqa: E402
from clip_decontamination_b200._lib import ATTN  # noqa: E402


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    dev = 'cuda'
    print(torch.cuda.get_device_name(0))
    for (M, N, K, tag) in [(3152, 2304, 768, 'qkv'), (3152, 768, 768, 'out'), (3152, 3072, 768, 'fc1'),
                           (3152, 768, 3072, 'fc2'), (16 * 50176, 512, 512, 'fixup512'), (16 * 50176, 128, 128, 'jbufix'),
                           (8192, 8192, 8192, 'square')]:
        A = torch.randn(M, K, device=dev).bfloat16()
        B = torch.randn(N, K, device=dev).bfloat16()
        C = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        ms = timeit(lambda: ops.gemm(A, B, C))
        ms_t = timeit(lambda: torch.matmul(A, B.t(), out=C))
        print(f'gemm {tag:9s} M={M} N={N} K={K}: {ms:.3f} ms = {2 * M * N * K / ms / 1e9:.1f} TFLOP/s '
              f'(torch/cuBLAS {ms_t:.3f} ms = {2 * M * N * K / ms_t / 1e9:.1f})')
        del A, B, C
    n, L, heads, hd = 16, 197, 12, 64
    qkv = torch.randn(n * L, 3 * heads * hd, device=dev).bfloat16()
    out = torch.empty(n * L, heads * hd, device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda: ops.attention(qkv, n, L, heads, hd, ATTN['STD'], out))
    print(f'attention std n={n}: {ms:.3f} ms')
    sim = torch.rand(n, L - 1, L - 1, device=dev)
    ms = timeit(lambda: ops.attention(qkv, n, L, heads, hd, ATTN['Experimental'], out, simmap=sim))
    print(f'attention experimental n={n}: {ms:.3f} ms')
    x = torch.randn(n * L, 768, device=dev)
    h = torch.empty(n * L, 768, device=dev, dtype=torch.bfloat16)
    g, b = torch.ones(768, device=dev), torch.zeros(768, device=dev)
    print(f'layernorm: {timeit(lambda: ops.layernorm(x, g, b, h)):.4f} ms')
    # JBU last stage, 16 crops
    C = 512
    src = torch.randn(n * 112 * 112, C, device=dev).bfloat16()
    kern = torch.rand(n * 224 * 224, 128, device=dev).bfloat16()
    dst = torch.empty(n * 224 * 224, C, device=dev, dtype=torch.bfloat16)
    hr = torch.empty_like(dst)
    ms = timeit(lambda: ops.jbu_apply(src, n, 112, 112, C, kern, 5, dst, hr), iters=3, warm=1)
    print(f'jbu_apply 112->224 n={n} C={C}: {ms:.3f} ms ({n * 224 * 224 * C * 121 * 2 / ms / 1e9:.1f} TFLOP/s fp32 FMA)')
    proj = torch.randn(n * 224 * 224, 32, device=dev).half()
    guid = torch.randn(n * 224 * 224, 4, device=dev)
    ms = timeit(lambda: ops.jbu_range_kernel(proj, guid, n, 224, 224, 5, 0.3, 1.0, kern), iters=3, warm=1)
    print(f'jbu_range_kernel 224 n={n}: {ms:.3f} ms')
    text = torch.randn(8, C, device=dev)
    lg = torch.empty(n, 8, 224 * 224, device=dev)
    ms = timeit(lambda: ops.norm_sim(dst, C, n, 224 * 224, C, text, lg), iters=5, warm=1)
    print(f'norm_sim n={n}: {ms:.3f} ms ({dst.numel() * 2 / ms / 1e6:.0f} GB/s)')


if __name__ == '__main__':
    main()
