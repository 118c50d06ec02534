"""Device timing of the attention kernels on the guidance shapes."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from perceptor_b200 import native, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=128)
    ap.add_argument("--t", type=int, default=257)
    ap.add_argument("--heads", type=int, default=16)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--persist", type=int, default=1, help="0: one tile per CTA (round-1 kernels), 1: persistent fwd+bwd, 2: persistent bwd only, 3: persistent fwd only")
    ap.add_argument("--split", type=int, default=1, help="0: four softmax warps per forward CTA, 1: eight (rows split over two warps)")
    args = ap.parse_args()
    native.lib().pcg_attn_set_persist(args.persist)
    native.lib().pcg_attn_set_split(args.split)
    dev = torch.device("cuda", 0)
    n, t, h = args.n, args.t, args.heads
    d = h * 64
    qkv = torch.randn(n * t, 3 * d, device=dev)
    qkv[:, :d] *= 0.125
    qkv = qkv.to(torch.bfloat16)
    d_out = torch.randn(n * t, d, device=dev).to(torch.bfloat16)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(4):
        out, lse = ops.attn_fwd(qkv, n, t, h)
        ops.attn_bwd(qkv, out, d_out, lse, n, t, h)
    tf, tb = 0.0, 0.0
    for _ in range(args.iters):
        flush.zero_()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        out, lse = ops.attn_fwd(qkv, n, t, h)
        e[1].record()
        ops.attn_bwd(qkv, out, d_out, lse, n, t, h)
        e[2].record()
        torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1])
        tb += e[1].elapsed_time(e[2])
    tf, tb = tf / args.iters, tb / args.iters
    fl = 4.0 * t * t * 64 * h * n
    print(f"persist={args.persist} split={args.split} n={n} T={t} heads={h}: fwd {tf * 1e3:.1f} us ({fl / tf / 1e9:.0f} TF/s)  bwd {tb * 1e3:.1f} us ({2 * fl / tb / 1e9:.0f} TF/s algorithmic)")


if __name__ == "__main__":
    main()
