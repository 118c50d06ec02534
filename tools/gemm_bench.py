"""Per-shape timing of the tcgen05 GEMM on the shapes of the guidance step (device time, CUDA events)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from perceptor_b200 import native, ops  # noqa: E402

MODES = {"bf16": native.GEMM_BF16, "act": native.GEMM_BIAS_ACT, "resid": native.GEMM_RESID_F32, "dact": native.GEMM_DACT,
         "f32": native.GEMM_F32}


def shapes(model, n_cut):
    if model == "l14":
        t, d, kpad, g2 = 257, 1024, 640, 256
    else:
        t, d, kpad, g2 = 50, 768, 3072, 49
    m, p = n_cut * t, n_cut * g2
    return [("patch", "f32", p, d, kpad), ("qkv", "bf16", m, 3 * d, d), ("out", "resid", m, d, d), ("fc", "act", m, 4 * d, d),
            ("proj", "resid", m, d, 4 * d), ("dproj", "dact", m, 4 * d, d), ("dfc", "bf16", m, d, 4 * d),
            ("dout", "bf16", m, d, d), ("dqkv", "bf16", m, d, 3 * d), ("dpatch", "bf16", p, kpad, d)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="l14")
    ap.add_argument("--cutouts", type=int, default=128)
    ap.add_argument("--bns", default="0,128,192,256")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--variant", type=int, default=1)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    native.lib().pcg_gemm_set_variant(args.variant)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    total = {}
    for name, mode, m, n, k in shapes(args.model, args.cutouts):
        if args.only and name not in args.only.split(","):
            continue
        a = torch.randn(m, k, device=dev).to(torch.bfloat16)
        b = (torch.randn(n, k, device=dev) / k**0.5).to(torch.bfloat16)
        bias = torch.randn(n, device=dev)
        aux = torch.randn(m, n, device=dev) if mode == "resid" else torch.randn(m, n, device=dev).to(torch.bfloat16)
        row = [f"{name:7s} {mode:5s} M={m:6d} N={n:5d} K={k:5d}"]
        for bn in [int(x) for x in args.bns.split(",")]:
            try:
                for _ in range(3):
                    ops.gemm(MODES[mode], a, b, bias=bias if mode != "dact" else None, aux=aux, bn=bn)
                ms = 0.0
                for _ in range(args.iters):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    ops.gemm(MODES[mode], a, b, bias=bias if mode != "dact" else None, aux=aux, bn=bn)
                    e1.record()
                    torch.cuda.synchronize()
                    ms += e0.elapsed_time(e1)
                ms /= args.iters
                tf = 2.0 * m * n * k / (ms * 1e-3) / 1e12
                row.append(f"bn{bn}: {ms * 1e3:7.1f}us {tf:6.0f}TF")
                total.setdefault(bn, 0.0)
                total[bn] += ms
            except Exception as e:  # noqa: BLE001
                row.append(f"bn{bn}: ERR {str(e)[:40]}")
        print("  ".join(row), flush=True)
    print("sum ms per bn:", {k: round(v, 3) for k, v in total.items()})


if __name__ == "__main__":
    main()
