"""Where does the bf16 path lose accuracy with outlier channels?  Native loss / gradient vs the fp32 CPU oracle on
ViT-B/32 with ~1 % of selected parameters scaled (tests/test_gpu_guidance.py::test_outlier_channels_match_oracle)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import guidance as guidance_oracle  # noqa: E402
from perceptor_b200 import cutouts, native  # noqa: E402
from perceptor_b200.guidance import GuidanceEngine, GuidanceLossFn  # noqa: E402
from perceptor_b200.vit import SHAPES, random_state_dict  # noqa: E402


def make_sd(shape, seed, factor, which):
    sd = random_state_dict(shape, seed)
    g = torch.Generator().manual_seed(seed + 7)
    for k in list(sd):
        v = sd[k]
        ln = (".ln_1.weight" in k or ".ln_2.weight" in k or k == "ln_pre.weight") and v.dim() == 1
        fc = k.endswith("mlp.c_fc.weight")
        if (ln and "ln" in which) or (fc and "fc" in which):
            n = v.shape[0]
            idx = torch.randperm(n, generator=g)[:max(1, n // 100)]
            v = v.clone()
            v[idx] *= factor
            sd[k] = v
    return sd


def cosine(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(a @ b / (a.norm() * b.norm() + 1e-300))


def main():
    dev = torch.device("cuda", 0)
    shape = SHAPES["ViT-B-32"]
    g = torch.Generator().manual_seed(31)
    images = torch.rand(1, 3, 224, 256, generator=g)
    rows = cutouts.sample_cutouts(torch.Generator().manual_seed(9), 1, 224, 256, 6, 1.0, 64, 224).tolist()
    targets = torch.nn.functional.normalize(torch.randn(2, shape.embed, generator=g))
    tw = torch.ones(2)
    torch.set_num_threads(os.cpu_count() or 1)
    for which in ("none", "ln", "fc", "ln+fc"):
        for factor in (10.0, 30.0, 100.0):
            if which == "none" and factor != 10.0:
                continue
            sd = make_sd(shape, 3, factor, which)
            img_ref = images.clone().requires_grad_()
            loss_ref = guidance_oracle.guidance_loss(img_ref, rows, sd, shape.image_size, shape.patch, shape.layers,
                                                     shape.heads, targets, tw, 1.0)
            loss_ref.backward()
            # the same fp32 oracle with bf16-rounded weights: what weight quantisation alone costs
            sd_q = {k: v.bfloat16().float() if v.dim() >= 2 else v for k, v in sd.items()}
            img_q = images.clone().requires_grad_()
            loss_q = guidance_oracle.guidance_loss(img_q, rows, sd_q, shape.image_size, shape.patch, shape.layers,
                                                   shape.heads, targets, tw, 1.0)
            loss_q.backward()
            eng = GuidanceEngine(shape, sd, dev, native.ACT_QUICKGELU)
            img = images.to(dev).requires_grad_()
            loss = GuidanceLossFn.apply(img, eng, eng.plan_cutouts(np.asarray(rows, dtype=np.int32)), targets.to(dev),
                                        tw.to(dev), 1.0, None)
            loss.backward()
            rel = abs(float(loss) - float(loss_ref)) / abs(float(loss_ref))
            print(f"{which:6s} x{factor:5.0f}: loss rel {rel:.2e}  grad cos {cosine(img.grad.cpu(), img_ref.grad):.6f}"
                  f"  | fp32 with bf16 weights: loss rel {abs(float(loss_q) - float(loss_ref)) / abs(float(loss_ref)):.2e}"
                  f"  grad cos {cosine(img_q.grad, img_ref.grad):.6f}  | grad norm ratio "
                  f"{float(img.grad.norm()) / float(img_ref.grad.norm()):.4f}", flush=True)
            del eng


if __name__ == "__main__":
    main()
