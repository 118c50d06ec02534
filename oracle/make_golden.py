"""Generate tests/golden/*.npz by running the REFERENCE's own in-tree files (imported by path from /root/reference).

Run in the build container only (`python -m oracle.make_golden`); the GPU box has no /root/reference, which is why
the outputs are committed.  The `perceptor` package itself cannot be imported offline (SURVEY.md §0 finding 4), so:
  * resize      = perceptor/transforms/resize/resize_right.py            (imported, unmodified)
  * ViT         = perceptor/models/ruclip/model.py VisionTransformer     (imported, unmodified)
  * composition = perceptor/models/open_clip.py:109-123 and perceptor/losses/clip/clip.py:89-99, written out here
                  with the reference's own expressions (those two files need open_clip / lantern to import).
"""
import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

REF = Path("/root/reference/perceptor")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"

TINY = dict(image_size=32, patch=8, width=128, layers=2, heads=2, embed=16)
MEAN = (0.48145466, 0.4578275, 0.40821073)  # perceptor/models/ruclip/processor.py:23-24
STD = (0.26862954, 0.26130258, 0.27577711)


def load_reference():
    sys.path.insert(0, str(REF / "transforms"))
    sys.path.insert(0, str(REF / "models" / "ruclip"))
    from resize import resize, resize_right  # noqa: E402
    import model as ruclip_model  # noqa: E402

    return resize, resize_right, ruclip_model


def tiny_state_dict(seed=7):
    from perceptor_b200.vit import VitShape, random_state_dict

    sd = random_state_dict(VitShape(**TINY), seed)
    g = torch.Generator().manual_seed(seed + 1)
    for k in list(sd):
        if sd[k].dim() == 1:  # LayerNorm affine + biases: make them non-trivial
            sd[k] = sd[k] + 0.1 * torch.randn(sd[k].shape, generator=g)
    return sd


def sd_checksums(sd):
    return np.array([[float(v.double().sum()), float((v.double() ** 2).sum())] for v in sd.values()])


def main():
    resize, rr, ruclip_model = load_reference()
    OUT.mkdir(parents=True, exist_ok=True)
    eps = float(torch.finfo(torch.float32).eps)

    # (ii) resize tap tables straight from the reference's helper functions ---------------------------------
    tables = {}
    for in_sz, out_sz in [(225, 224), (256, 224), (300, 224), (512, 224), (100, 224), (768, 336), (400, 336), (200, 336),
                          (37, 32), (20, 32)]:
        scale = float(out_sz / in_sz)
        method = rr.interpolation_methods.methods["lanczos3" if in_sz >= out_sz else "bicubic"]
        grid = rr.get_projected_grid(in_sz, out_sz, scale, torch, False, None)
        cur_method, cur_support = rr.apply_antialiasing_if_needed(method, method.support_sz, scale, True)
        fov = rr.get_field_of_view(grid, cur_support, torch, eps, None)
        left_unpadded = fov[:, 0].clone()
        pad_sz, grid, fov = rr.calc_pad_sz(in_sz, out_sz, fov, grid, scale, False, torch, None)
        weights = rr.get_weights(cur_method, grid, fov)
        tables[f"left_{in_sz}_{out_sz}"] = left_unpadded.numpy().astype(np.int32)
        tables[f"w_{in_sz}_{out_sz}"] = weights.numpy().astype(np.float32)
    np.savez_compressed(OUT / "resize_tables.npz", **tables)

    # (ii') resize outputs + input gradients on small seeded inputs -------------------------------------------
    g = torch.Generator().manual_seed(0)
    small = {}
    for i, (h, w, oh, ow) in enumerate([(40, 52, 16, 16), (20, 12, 32, 32), (37, 37, 32, 32), (50, 20, 32, 32),
                                        (32, 32, 32, 32), (33, 64, 32, 32), (9, 70, 16, 24)]):
        x = torch.rand(2, 3, h, w, generator=g).requires_grad_()
        y = resize(x, out_shape=(oh, ow))
        cot = torch.randn(y.shape, generator=g)
        (gx,) = torch.autograd.grad(y, x, cot)
        small[f"x{i}"], small[f"y{i}"] = x.detach().numpy(), y.detach().numpy()
        small[f"cot{i}"], small[f"gx{i}"] = cot.numpy(), gx.numpy()
    np.savez_compressed(OUT / "resize_small.npz", **small)

    # (iii) tiny ViT through the reference's VisionTransformer ---------------------------------------------
    sd = tiny_state_dict()
    vit = ruclip_model.VisionTransformer(TINY["image_size"], TINY["patch"], TINY["width"], TINY["layers"],
                                         TINY["heads"], TINY["embed"]).eval().requires_grad_(False)
    vit.load_state_dict(sd)
    x = torch.randn(3, 3, 32, 32, generator=g).requires_grad_()
    enc = vit(x)
    cot = torch.randn(enc.shape, generator=g)
    (gx,) = torch.autograd.grad(enc, x, cot)
    np.savez_compressed(OUT / "vit_tiny.npz", x=x.detach().numpy(), enc=enc.detach().numpy(), cot=cot.numpy(),
                        gx=gx.numpy(), sd_checksums=sd_checksums(sd))

    # (iv) whole path on the tiny config: cutouts -> reference resize -> Normalize -> reference ViT ->
    #      F.normalize -> CLIP.forward's expression -> backward to the image -------------------------------
    images = torch.rand(2, 3, 48, 40, generator=g).requires_grad_()
    rows = [(0, 0, 0, 40), (1, 8, 0, 32), (0, 3, 5, 20), (1, 10, 2, 37), (0, 16, 8, 32)]
    targets = F.normalize(torch.randn(3, TINY["embed"], generator=g))  # clip.py:72 re-normalises targets
    weights = torch.tensor([1.0, 0.5, -0.25])
    multiplier = 0.01
    mean = torch.tensor(MEAN).reshape(1, 3, 1, 1)
    std = torch.tensor(STD).reshape(1, 3, 1, 1)
    pixels = torch.cat([
        (resize(images[b:b + 1, :, y0:y0 + s, x0:x0 + s], out_shape=(32, 32)) - mean) / std  # open_clip.py:111-118
        for b, y0, x0, s in rows])
    image_encodings = F.normalize(vit(pixels))  # open_clip.py:120-121
    spherical_distance = (image_encodings[:, None] - targets[None, :]).norm(dim=2).div(2).arcsin().square().mul(2)
    loss = (spherical_distance * weights).mean().mul(multiplier)  # clip.py:91-99
    (gi,) = torch.autograd.grad(loss, images)
    np.savez_compressed(OUT / "guidance_tiny.npz", images=images.detach().numpy(), rows=np.array(rows, dtype=np.int32),
                        targets=targets.numpy(), weights=weights.numpy(), multiplier=multiplier,
                        pixels=pixels.detach().numpy(), encodings=image_encodings.detach().numpy(),
                        loss=float(loss), grad=gi.numpy(), sd_checksums=sd_checksums(sd))

    # (i) cutout rows: no referent in the reference; pins the spec against drift (oracle restatement) ---------
    from oracle import sampler as sampler_oracle

    cases = {}
    for j, (seed, b, h, w, n, pw, lo, hi) in enumerate([(0, 1, 256, 256, 16, 1.0, 64, 256), (1, 4, 512, 512, 64, 1.0, 64, 512),
                                                        (2, 1, 768, 768, 256, 0.5, 100, 768), (3, 2, 300, 500, 5, 2.0, 32, 300)]):
        rows_j = sampler_oracle.sample_cutouts(torch.Generator().manual_seed(seed), b, h, w, n, pw, lo, hi)
        cases[f"args{j}"] = np.array([seed, b, h, w, n, lo, hi], dtype=np.int64)
        cases[f"pow{j}"] = np.array(pw)
        cases[f"rows{j}"] = np.array(rows_j, dtype=np.int32)
    np.savez_compressed(OUT / "sampler_rows.npz", **cases)
    for p in sorted(OUT.glob("*.npz")):
        print(p.name, p.stat().st_size)


if __name__ == "__main__":
    main()
