"""Phase timeline of the tcgen05 attention kernels: per-CTA clock64 stamps (pcg_attn_set_trace), median over CTAs.

Also prints per-iteration device times so that outliers are visible."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from perceptor_b200 import native, ops  # noqa: E402

# persistent forward: stamps of each CTA's SECOND item (steady state), warp 0
FWD2 = {0: "item start", 9: "next item start", 4: "S ready", 5: "max pass done", 6: "exp pass done", 7: "O ready", 8: "epilogue done",
        17: "grp1: S ready", 18: "grp1: exp pass done", 19: "grp1: epilogue done",
        10: "ctl: Q+K landed", 11: "ctl: tmem free (S issue)", 12: "ctl: next Q+K requested", 13: "ctl: PV hi issued",
        14: "ctl: P lo seen", 15: "ctl: O retired", 16: "ctl: next V requested", 20: "edge: vectors written",
        21: "edge: K landed", 22: "edge: K dots done", 23: "edge: row done"}
FWD = {0: "start", 1: "setup done", 2: "K+Q landed", 3: "V landed", 4: "S ready (w0)", 5: "max pass done", 6: "exp pass done",
       7: "O ready", 8: "epilogue done", 9: "edge row done", 10: "all warps done"}
# persistent backward: warp 0's stamps of one steady-state item
BWD2 = {0: "item start (vectors wait)", 1: "b0 S/dP ready", 2: "b0 alu done", 3: "b1 S/dP ready", 4: "b1 alu done",
        5: "b2 S/dP ready", 6: "b2 alu done", 7: "b3 S/dP ready", 8: "b3 alu done", 9: "next item start",
        10: "edge: item start", 11: "edge: tiles landed", 13: "edge: dots done", 12: "edge: gemv done (EdgeDone)",
        16: "drain: next delta prepared", 14: "drain: tile0 complete seen", 15: "drain: dV0/dK0 drained",
        18: "drain: tile1 complete seen", 19: "drain: dV1/dK1 drained", 17: "drain: dQ drained", 20: "ctl: b0 half1 seen", 21: "ctl: b1 half1 seen",
        22: "ctl: b2 half1 seen", 23: "ctl: b3 half1 seen", 24: "ctl: b0 grads issue", 25: "ctl: b1 grads issue",
        26: "ctl: b2 grads issue", 27: "ctl: b3 grads issue"}
BWD = {0: "start", 1: "first loads landed", 2: "edge vectors ready", 3: "edge gemv done", 4: "b0 S/dP ready", 5: "b0 alu done",
       6: "b1 S/dP ready", 7: "b1 alu done", 8: "b2 S/dP ready", 9: "b2 alu done", 10: "b3 S/dP ready", 11: "b3 alu done",
       12: "tile0 mma done", 13: "tile1 mma done", 14: "epilogues done", 15: "all warps done", 16: "ctl b0 P/dS seen",
       17: "ctl b1 P/dS seen", 18: "ctl b2 P/dS seen", 19: "ctl b3 P/dS seen", 20: "delta done", 21: "ctl issued first S/dP"}


def show(trace, names, title, flag=None):
    t = trace.cpu().double()
    valid = t[:, 0] > 0
    if flag is not None:
        valid &= t[:, 28] == flag
    t = t[valid]
    if t.shape[0] == 0:
        return
    rel = t - t[:, :1]
    print(f"--- {title}: {t.shape[0]} CTAs, cycles since CTA start (median / p90)")
    order = sorted(names, key=lambda k: float(rel[:, k][t[:, k] > 0].median()) if (t[:, k] > 0).any() else 1e18)
    for k in order:
        m = t[:, k] > 0
        if not m.any():
            continue
        v = rel[:, k][m]
        print(f"  {names[k]:22s} {v.median():9.0f} {v.quantile(0.9):9.0f}   ({int(m.sum())} CTAs)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=128)
    ap.add_argument("--t", type=int, default=257)
    ap.add_argument("--heads", type=int, default=16)
    ap.add_argument("--iters", type=int, default=8)
    ap.add_argument("--persist", type=int, default=1, help="0: one tile per CTA (round-1 kernels), 1: persistent fwd+bwd, 2: persistent bwd only, 3: persistent fwd only")
    ap.add_argument("--split", type=int, default=1)
    ap.add_argument("--no-bwd", action="store_true")
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
    for _ in range(3):
        out, lse = ops.attn_fwd(qkv, n, t, h)
        ops.attn_bwd(qkv, out, d_out, lse, n, t, h)
    tf, tb = [], []
    for _ in range(args.iters):
        flush.zero_()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        out, lse = ops.attn_fwd(qkv, n, t, h)
        e[1].record()
        ops.attn_bwd(qkv, out, d_out, lse, n, t, h)
        e[2].record()
        torch.cuda.synchronize()
        tf.append(e[0].elapsed_time(e[1]) * 1e3)
        tb.append(e[1].elapsed_time(e[2]) * 1e3)
    print("fwd us per iteration:", " ".join(f"{x:.0f}" for x in tf))
    print("bwd us per iteration:", " ".join(f"{x:.0f}" for x in tb))
    n_cta = 2 * h * n
    trace = torch.zeros(n_cta, 32, dtype=torch.int64, device=dev)
    native.lib().pcg_attn_set_trace(trace.data_ptr())
    out, lse = ops.attn_fwd(qkv, n, t, h)
    torch.cuda.synchronize()
    if args.persist in (1, 3):
        tt = trace.cpu().double()
        tt = tt[tt[:, 30] > 0]
        life = tt[:, 31] - tt[:, 30]
        print(f"persistent forward: {tt.shape[0]} CTAs, items per CTA {tt[:, 29].min():.0f}-{tt[:, 29].max():.0f}, CTA lifetime "
              f"median {life.median():.0f} max {life.max():.0f} clk, per item {float((life / tt[:, 29]).median()):.0f} clk")
    if args.persist in (1, 3) and args.split and t >= 130:
        show(trace, FWD2, "forward (split rows), traced item WITHOUT the edge query row", flag=1)
        show(trace, FWD2, "forward (split rows), traced item WITH the edge query row", flag=2)
    else:
        show(trace, FWD2 if args.persist in (1, 3) else FWD,
             "forward (persistent, one steady-state item of each CTA)" if args.persist in (1, 3) else "forward")
    if args.no_bwd:
        native.lib().pcg_attn_set_trace(None)
        return
    trace.zero_()
    ops.attn_bwd(qkv, out, d_out, lse, n, t, h)
    torch.cuda.synchronize()
    native.lib().pcg_attn_set_trace(None)
    if args.persist in (1, 2) and t >= 130:
        tt = trace.cpu().double()
        tt = tt[tt[:, 30] > 0]
        life = tt[:, 31] - tt[:, 30]
        print(f"persistent backward: {tt.shape[0]} CTAs, items per CTA {tt[:, 29].min():.0f}-{tt[:, 29].max():.0f}, CTA lifetime "
              f"median {life.median():.0f} max {life.max():.0f} clk, per item {float((life / tt[:, 29]).median()):.0f} clk")
        show(trace, BWD2, "backward (persistent, one steady-state item of each CTA)")
    else:
        show(trace[: h * n], BWD, "backward")


if __name__ == "__main__":
    main()
