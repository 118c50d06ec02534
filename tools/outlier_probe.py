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


def perturb(sd, seed=1):
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in sd.items():
        if k.endswith("weight") and v.dim() == 1:
            out[k] = v + 0.1 * torch.randn(v.shape, generator=g)
        elif k.endswith("bias") or k.endswith("in_proj_bias"):
            out[k] = v + 0.02 * torch.randn(v.shape, generator=g)
        else:
            out[k] = v
    return out


def make_sd(shape, seed, factor, which):
    sd = perturb(random_state_dict(shape, seed))
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


def autocast_reference(images, rows, sd, shape, targets, tw, dev, dtype):
    """The SAME network in stock PyTorch on the GPU under torch.autocast(dtype) -- what the reference's own reduced-
    precision path does (perceptor/models/open_clip.py:109 runs encode_images under autocast): eager cuBLAS bf16 / fp16
    GEMMs, fp32 LayerNorm / softmax, fp32 weights.  Resize + normalise stay fp32 on the CPU (oracle)."""
    import torch.nn.functional as F

    from oracle import loss as loss_oracle
    from oracle import vit as vit_oracle

    img = images.clone().requires_grad_()
    pixels = guidance_oracle.cutout_pixels(img, rows, shape.image_size).to(dev)
    sdd = {k: v.to(dev) for k, v in sd.items()}
    d = sdd["conv1.weight"].shape[0]
    with torch.autocast("cuda", dtype=dtype):
        x = F.conv2d(pixels, sdd["conv1.weight"], stride=shape.patch)
        x = x.reshape(x.shape[0], d, -1).permute(0, 2, 1)
        cls = sdd["class_embedding"].to(x.dtype) + torch.zeros(x.shape[0], 1, d, dtype=x.dtype, device=dev)
        x = torch.cat([cls, x], dim=1) + sdd["positional_embedding"]
        x = F.layer_norm(x, (d,), sdd["ln_pre.weight"], sdd["ln_pre.bias"], 1e-5)
        for i in range(shape.layers):
            x = vit_oracle.block(x, sdd, f"transformer.resblocks.{i}.", shape.heads, "quickgelu")
        x = F.layer_norm(x[:, 0, :], (d,), sdd["ln_post.weight"], sdd["ln_post.bias"], 1e-5)
        enc = x @ sdd["proj"]
    enc = F.normalize(enc.float())
    loss = loss_oracle.clip_loss(enc, targets.to(dev), tw.to(dev), 1.0)
    loss.backward()
    return float(loss), img.grad


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
        for factor in (10.0, 30.0):
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
            loss_ac, grad_ac = autocast_reference(images, rows, sd, shape, targets, tw, dev, torch.bfloat16)
            loss_h, grad_h = autocast_reference(images, rows, sd, shape, targets, tw, dev, torch.float16)
            rel = abs(float(loss) - float(loss_ref)) / abs(float(loss_ref))
            print(f"{which:6s} x{factor:5.0f}: loss rel {rel:.2e}  grad cos {cosine(img.grad.cpu(), img_ref.grad):.6f}"
                  f"  | fp32 with bf16 weights: loss rel {abs(float(loss_q) - float(loss_ref)) / abs(float(loss_ref)):.2e}"
                  f"  grad cos {cosine(img_q.grad, img_ref.grad):.6f}  | torch autocast bf16: loss rel "
                  f"{abs(loss_ac - float(loss_ref)) / abs(float(loss_ref)):.2e} grad cos {cosine(grad_ac, img_ref.grad):.6f}"
                  f"  | torch autocast fp16: grad cos {cosine(grad_h, img_ref.grad):.6f}  | grad norm ratio "
                  f"{float(img.grad.norm()) / float(img_ref.grad.norm()):.4f}", flush=True)
            del eng


if __name__ == "__main__":
    main()
