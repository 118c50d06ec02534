"""Small multi-item run of the persistent attention kernels for compute-sanitizer (one tool per gpurun call):
    compute-sanitizer --tool memcheck python tools/attn_memcheck.py
10 x 16 heads x 2 tiles = 320 forward items on 296 CTAs, 160 backward items on 148: some CTAs walk two items."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from perceptor_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
for n, t, h in ((10, 257, 16), (14, 197, 12)):
    d = h * 64
    g = torch.Generator().manual_seed(0)
    qkv = torch.randn(n * t, 3 * d, generator=g)
    qkv[:, :d] *= 0.25
    qkv = qkv.to(dev, torch.bfloat16)
    d_out = torch.randn(n * t, d, generator=g).to(dev, torch.bfloat16)
    out, lse = ops.attn_fwd(qkv, n, t, h)
    d_qkv = ops.attn_bwd(qkv, out, d_out, lse, n, t, h)
    torch.cuda.synchronize()
    print(n, t, h, float(out.float().abs().mean()), float(d_qkv.float().abs().mean()), bool(torch.isfinite(d_qkv.float()).all()))
