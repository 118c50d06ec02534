"""Diagnostic: how much of the guidance step is launch gaps?  Times the same forward + backward (fixed cutout table)
launched eagerly and replayed from one CUDA graph."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from perceptor_b200 import cutouts, native  # noqa: E402
from perceptor_b200.guidance import GuidanceEngine, GuidanceLossFn  # noqa: E402
from perceptor_b200.vit import SHAPES, random_state_dict  # noqa: E402


def main():
    arch, hw, n_cut = (sys.argv[1] if len(sys.argv) > 1 else "ViT-L-14"), 512, int(sys.argv[2]) if len(sys.argv) > 2 else 128
    dev = torch.device("cuda", 0)
    shape = SHAPES[arch]
    eng = GuidanceEngine(shape, random_state_dict(shape, 0), dev, native.ACT_QUICKGELU)
    g = torch.Generator().manual_seed(0)
    images = torch.rand(1, 3, hw, hw, generator=g).to(dev).requires_grad_()
    targets = torch.nn.functional.normalize(torch.randn(2, shape.embed, generator=g)).to(dev)
    tw = torch.ones(2, device=dev)
    rows = cutouts.sample_cutouts(torch.Generator().manual_seed(0), 1, hw, hw, n_cut, 1.0, 128, hw)
    plan = eng.plan_cutouts(np.asarray(rows))

    def step():
        loss = GuidanceLossFn.apply(images, eng, plan, targets, tw, 1.0, None)
        (grad,) = torch.autograd.grad(loss, images)
        return loss, grad

    def timed(fn, k=10):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / k

    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.synchronize()
    eager = timed(step)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        loss, grad = step()
    graph.replay()
    torch.cuda.synchronize()
    ref_loss, ref_grad = step()
    print(f"graph vs eager: loss {float(loss.detach()):.6f} / {float(ref_loss.detach()):.6f}, "
          f"grad max diff {float((grad - ref_grad).abs().max()):.3e}")
    print(f"{arch} x {n_cut}: eager {eager:.3f} ms/step, graph replay {timed(graph.replay):.3f} ms/step, eager again {timed(step):.3f}")


if __name__ == "__main__":
    main()
