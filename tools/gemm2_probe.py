"""Probe of the CTA-pair (cta_group::2) GEMM: the four big ViT shapes against a torch fp32 reference on the same bf16 operands,
with CUDA-event timings.  CSEG_GEMM_2CTA=0 selects the single-CTA kernel for an A/B run."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from clip_decontamination_b200 import ops  # noqa: E402
from clip_decontamination_b200._lib import ACT_GELU, ACT_NONE  # noqa: E402

torch.manual_seed(0)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 18912
cases = [('fc1-noact', 3072, 768, None, False, torch.bfloat16), ('fc1-f32out', 3072, 768, None, False, torch.float32),
         ('qkv', 2304, 768, None, False, torch.bfloat16), ('out', 768, 768, None, True, torch.float32),
         ('fc1', 3072, 768, ACT_GELU, False, torch.bfloat16), ('fc2', 768, 3072, None, True, torch.float32)]
for name, N, K, act, res, odt in cases:
    A = (torch.randn(M, K, device='cuda') * 0.5).to(torch.bfloat16)
    B = (torch.randn(N, K, device='cuda') * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device='cuda')
    x = torch.randn(M, N, device='cuda') if res else None
    ref = A.float() @ B.float().t() + bias
    if act == ACT_GELU:
        ref = torch.nn.functional.gelu(ref)
    if res:
        ref = ref + x
    out = x.clone() if res else torch.empty(M, N, device='cuda', dtype=odt)
    kw = dict(bias=bias)
    if act is not None:
        kw['act'] = act
    if res:
        kw['residual'] = out
    ops.gemm(A, B, out, **kw)
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    # timing (the residual case accumulates in place: values drift, timing only)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        ops.gemm(A, B, out, **kw)
    s.record()
    for _ in range(20):
        ops.gemm(A, B, out, **kw)
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) / 20 * 1e3
    print(f'{name}: M={M} N={N} K={K} max|err|={err:.3e} (|ref|max {scale:.2f})  {us:.1f} us  {2.0 * M * N * K / us / 1e6:.0f} TFLOP/s', flush=True)
