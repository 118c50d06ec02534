"""Pin the CPU oracle against vectors produced by the reference's own in-tree files (oracle/make_golden.py)."""
from pathlib import Path

import numpy as np
import torch

from oracle import guidance as guidance_oracle
from oracle import loss as loss_oracle
from oracle import resize as resize_oracle
from oracle import sampler as sampler_oracle
from oracle import vit as vit_oracle
from oracle.make_golden import TINY, sd_checksums, tiny_state_dict

GOLDEN = Path(__file__).parent / "golden"


def test_resize_taps_match_reference_tables():
    z = np.load(GOLDEN / "resize_tables.npz")
    pairs = sorted({tuple(int(v) for v in k.split("_")[1:]) for k in z.files})
    assert len(pairs) == 10
    for in_sz, out_sz in pairs:
        fov, w = resize_oracle.dim_taps(in_sz, out_sz, "lanczos3" if in_sz >= out_sz else "cubic")
        assert np.array_equal(fov[:, 0].numpy().astype(np.int32), z[f"left_{in_sz}_{out_sz}"]), (in_sz, out_sz)
        assert w.shape == z[f"w_{in_sz}_{out_sz}"].shape
        assert np.array_equal(w.numpy(), z[f"w_{in_sz}_{out_sz}"]), (in_sz, out_sz)  # same fp32 op sequence
    # tap counts quoted in SURVEY.md §8(c)
    taps = {(i, o): z[f"w_{i}_{o}"].shape[1] for i, o in pairs}
    assert [taps[p] for p in [(225, 224), (256, 224), (300, 224), (512, 224), (100, 224), (768, 336), (400, 336),
                              (200, 336)]] == [7, 7, 9, 14, 4, 14, 8, 4]


def test_resize_outputs_and_gradients_match_reference():
    z = np.load(GOLDEN / "resize_small.npz")
    n = len([k for k in z.files if k.startswith("x")])
    assert n == 7
    for i in range(n):
        x = torch.from_numpy(z[f"x{i}"]).requires_grad_()
        want = torch.from_numpy(z[f"y{i}"])
        y = resize_oracle.resize(x, want.shape[-2:])
        assert y.shape == want.shape
        assert float((y - want).abs().max()) <= 1e-6, i
        (gx,) = torch.autograd.grad(y, x, torch.from_numpy(z[f"cot{i}"]))
        assert float((gx - torch.from_numpy(z[f"gx{i}"])).abs().max()) <= 1e-5, i


def test_vit_matches_reference_vision_transformer():
    z = np.load(GOLDEN / "vit_tiny.npz")
    sd = tiny_state_dict()
    assert np.allclose(sd_checksums(sd), z["sd_checksums"], rtol=1e-12), "weight generator drifted; regenerate golden"
    x = torch.from_numpy(z["x"]).requires_grad_()
    enc = vit_oracle.encode(x, sd, TINY["patch"], TINY["layers"], TINY["heads"])
    assert float((enc - torch.from_numpy(z["enc"])).abs().max()) <= 2e-5
    (gx,) = torch.autograd.grad(enc, x, torch.from_numpy(z["cot"]))
    want = torch.from_numpy(z["gx"])
    assert float((gx - want).norm() / want.norm()) <= 1e-4


def test_whole_path_matches_reference_composition():
    z = np.load(GOLDEN / "guidance_tiny.npz")
    sd = tiny_state_dict()
    images = torch.from_numpy(z["images"]).requires_grad_()
    rows = z["rows"].tolist()
    pixels = guidance_oracle.cutout_pixels(images, rows, TINY["image_size"])
    assert float((pixels - torch.from_numpy(z["pixels"])).abs().max()) <= 2e-6
    enc = guidance_oracle.encode_cutouts(images, rows, sd, TINY["image_size"], TINY["patch"], TINY["layers"], TINY["heads"])
    assert float((enc - torch.from_numpy(z["encodings"])).abs().max()) <= 2e-5
    loss = guidance_oracle.guidance_loss(images, rows, sd, TINY["image_size"], TINY["patch"], TINY["layers"],
                                         TINY["heads"], torch.from_numpy(z["targets"]), torch.from_numpy(z["weights"]),
                                         float(z["multiplier"]))
    assert abs(float(loss) - float(z["loss"])) <= 1e-5 * abs(float(z["loss"]))
    (g,) = torch.autograd.grad(loss, images)
    want = torch.from_numpy(z["grad"])
    assert float((g - want).norm() / want.norm()) <= 1e-3


def test_loss_formula_edge_cases():
    e = torch.nn.functional.normalize(torch.randn(3, 8, generator=torch.Generator().manual_seed(0)))
    d = loss_oracle.spherical_distance(e, e)
    assert float(d.diagonal().abs().max()) == 0.0
    anti = loss_oracle.spherical_distance(e[:1], -e[:1])
    assert abs(float(anti) - 2 * (np.pi / 2) ** 2) <= 1e-2  # r = 2 -> theta = pi/2 (asin is ill-conditioned there)
    assert loss_oracle.default_multiplier("ViT-L-14") == 0.01 and loss_oracle.default_multiplier("ViT-B-32") == 1.0
    w = torch.tensor([1.0, -1.0, 0.0])
    assert abs(float(loss_oracle.clip_loss(e, e, w))) < 10


def test_sampler_rows_match_golden():
    z = np.load(GOLDEN / "sampler_rows.npz")
    for j in range(4):
        seed, b, h, w, n, lo, hi = (int(v) for v in z[f"args{j}"])
        rows = sampler_oracle.sample_cutouts(torch.Generator().manual_seed(seed), b, h, w, n, float(z[f"pow{j}"]), lo, hi)
        assert np.array_equal(np.array(rows, dtype=np.int32), z[f"rows{j}"])


def test_diffusion_glue_matches_reference_predictions():
    """oracle/diffusion.py against vectors from the reference's own predictions.py / velocity_diffusion.py /
    clamp_with_grad.py (oracle/make_golden_diffusion.py)."""
    from oracle import diffusion as dz

    z = np.load(GOLDEN / "diffusion_glue.npz")
    t = {k: torch.from_numpy(z[k]) for k in z.files}
    assert np.array_equal(dz.schedule_ts(50).numpy(), z["schedule_50"])
    assert np.array_equal(dz.schedule_ts(7, 0.9, 0.05, 5.0).numpy(), z["schedule_7"])
    args = (t["images"], t["ts"], t["velocities"])
    for name, got in [("denoised_xs", dz.denoised_xs(*args)), ("predicted_noise", dz.predicted_noise(*args)),
                      ("denoised_images", dz.denoised_images(*args)), ("step", dz.step(*args, t["to_ts"])),
                      ("guided", dz.guided(*args, t["guiding"], 0.7, 1e-6)),
                      ("forced", dz.forced_denoised_images(*args, t["images"].flip(0)))]:
        assert torch.equal(got, t[name]), name  # the same float32 expressions: bit-equal
    v, x = t["velocities"].clone().requires_grad_(), t["images"].clone().requires_grad_()
    gv, gx = torch.autograd.grad((dz.denoised_images(x, t["ts"], v) * t["cot"]).sum(), (v, x))
    assert torch.equal(gv, t["grad_velocities"]) and torch.equal(gx, t["grad_images"])
    one = (t["images"][:1], t["ts"][:1], t["velocities"][:1] * 3)
    assert torch.allclose(dz.dynamic_threshold_velocities(*one, 0.9), t["dynamic_threshold"], atol=1e-6)
    static = dz.forced_denoised_images(*one, dz.clamp_with_grad(dz.denoised_images(*one), 0, 1))
    assert torch.allclose(static, t["static_threshold"], atol=1e-6)
    zz = t["cwg_x"].clone().requires_grad_()
    y = dz.clamp_with_grad(zz, 0.0, 1.0)
    (gz,) = torch.autograd.grad(y, zz, t["cwg_gin"])
    assert torch.equal(y.detach(), t["cwg_y"]) and torch.equal(gz, t["cwg_gx"])
