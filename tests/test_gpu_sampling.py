"""Guided-sampling glue on the GPU: the fused native maps against the CPU oracle (which is pinned
to the reference's predictions.py by tests/golden/diffusion_glue.npz), and one guided step end to end."""
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import diffusion as dz  # noqa: E402
from oracle import guidance as guidance_oracle  # noqa: E402
from perceptor_b200 import native, transforms, velocity_diffusion as vd  # noqa: E402
from perceptor_b200.guidance import GuidanceEngine, GuidanceLossFn  # noqa: E402
from perceptor_b200.vit import VitShape, random_state_dict  # noqa: E402

GOLDEN = Path(__file__).parent / "golden"
TOL = 2e-6


def _close(a, b, tol=TOL):
    return float((a.detach().cpu() - b).abs().max()) <= tol


@pytest.mark.parametrize("shape", [(3, 3, 10, 12), (2, 3, 7, 9), (1, 3, 256, 256)])  # 7x9: the scalar (unaligned) kernel
def test_predictions_maps_match_oracle(cuda_device, shape):
    g = torch.Generator().manual_seed(sum(shape))
    images, velocities = torch.rand(shape, generator=g), torch.randn(shape, generator=g)
    ts = torch.rand(shape[0], generator=g) * 0.9 + 0.05
    to_ts = ts * 0.8
    guiding = torch.randn(shape, generator=g) * 2e-6
    p = vd.Predictions(images.to(cuda_device), ts, velocities.to(cuda_device))
    args = (images, ts, velocities)
    assert _close(p.denoised_xs, dz.denoised_xs(*args))
    assert _close(p.predicted_noise, dz.predicted_noise(*args))
    assert _close(p.denoised_images, dz.denoised_images(*args))
    assert _close(p.step(to_ts), dz.step(*args, to_ts))
    assert _close(p.guided(guiding.to(cuda_device), 0.7, 1e-6).velocities, dz.guided(*args, guiding, 0.7, 1e-6))
    assert _close(p.forced_denoised_images(images.flip(0).to(cuda_device)).velocities,
                  dz.forced_denoised_images(*args, images.flip(0)), 2e-5)
    # timesteps living on the device take the same path
    p_dev = vd.Predictions(images.to(cuda_device), ts.to(cuda_device), velocities.to(cuda_device))
    assert _close(p_dev.denoised_images, dz.denoised_images(*args))
    # autograd through the fused map
    v = velocities.to(cuda_device).requires_grad_()
    x = images.to(cuda_device).requires_grad_()
    cot = torch.randn(shape, generator=g)
    gv, gx = torch.autograd.grad((vd.Predictions(x, ts, v).denoised_images * cot.to(cuda_device)).sum(), (v, x))
    vr, xr = velocities.clone().requires_grad_(), images.clone().requires_grad_()
    gvr, gxr = torch.autograd.grad((dz.denoised_images(xr, ts, vr) * cot).sum(), (vr, xr))
    assert _close(gv, gvr) and _close(gx, gxr)


def test_golden_vectors_through_the_native_maps(cuda_device):
    z = np.load(GOLDEN / "diffusion_glue.npz")
    t = {k: torch.from_numpy(z[k]) for k in z.files}
    p = vd.Predictions(t["images"].to(cuda_device), t["ts"], t["velocities"].to(cuda_device))
    for name, got in [("denoised_xs", p.denoised_xs), ("predicted_noise", p.predicted_noise),
                      ("denoised_images", p.denoised_images), ("step", p.step(t["to_ts"])),
                      ("guided", p.guided(t["guiding"].to(cuda_device), 0.7, 1e-6).velocities)]:
        assert _close(got, t[name]), name
    p1 = vd.Predictions(t["images"][:1].to(cuda_device), t["ts"][:1], (t["velocities"][:1] * 3).to(cuda_device))
    assert _close(p1.dynamic_threshold(0.9).velocities, t["dynamic_threshold"], 2e-5)
    assert _close(p1.static_threshold().velocities, t["static_threshold"], 2e-5)
    x = t["cwg_x"].to(cuda_device).requires_grad_()
    y = transforms.clamp_with_grad(x, 0.0, 1.0)
    (gx,) = torch.autograd.grad(y, x, t["cwg_gin"].to(cuda_device))
    assert torch.equal(y.detach().cpu(), t["cwg_y"]) and torch.equal(gx.cpu(), t["cwg_gx"])


def test_step_with_eta_adds_the_right_amount_of_noise(cuda_device):
    g = torch.Generator().manual_seed(5)
    images, velocities = torch.rand(2, 3, 64, 64, generator=g), torch.randn(2, 3, 64, 64, generator=g)
    ts, to_ts, eta = torch.tensor([0.8, 0.6]), torch.tensor([0.7, 0.5]), 0.5
    p = vd.Predictions(images.to(cuda_device), ts, velocities.to(cuda_device))
    torch.manual_seed(0)
    out = p.step(to_ts, eta=eta).cpu()
    # predictions.py:82-99 with the noise removed: the deterministic part must match, the residual must be noise
    a, s = dz.t_to_alpha_sigma(ts)
    ta, tsg = dz.t_to_alpha_sigma(to_ts)
    ddim = eta * (tsg**2 / s**2).sqrt() * (1 - a**2 / ta**2).sqrt()
    adj = (tsg**2 - ddim**2).sqrt()
    base = dz.decode(dz.denoised_xs(images, ts, velocities) * ta[:, None, None, None]
                     + dz.predicted_noise(images, ts, velocities) * adj[:, None, None, None])
    resid = (out - base) * 2
    for i in range(2):
        assert abs(float(resid[i].std()) / float(ddim[i]) - 1.0) < 0.05 and abs(float(resid[i].mean())) < 0.02


def test_guided_step_matches_oracle_composition(cuda_device):
    """One CLIP-guided step: native loss on the native denoised images, gradient w.r.t. the velocities, guided(),
    step() -- against the same composition of the CPU oracles (fp32)."""
    shape = VitShape(image_size=32, patch=8, width=128, layers=2, heads=2, embed=16)
    sd = random_state_dict(shape, 3)
    g = torch.Generator().manual_seed(9)
    images, velocities = torch.rand(1, 3, 48, 48, generator=g), torch.randn(1, 3, 48, 48, generator=g) * 0.3
    targets = torch.nn.functional.normalize(torch.randn(2, shape.embed, generator=g))
    tw = torch.ones(2)
    rows = [(0, 0, 0, 48), (0, 5, 9, 32), (0, 16, 2, 30), (0, 1, 1, 40)]
    ts, to_ts, scale, clampv = torch.tensor([0.6]), torch.tensor([0.5]), 0.5, 1e-2

    eng = GuidanceEngine(shape, sd, cuda_device, native.ACT_QUICKGELU)
    plan = eng.plan_cutouts(np.asarray(rows, dtype=np.int32))

    def loss_fn(den):
        return GuidanceLossFn.apply(den, eng, plan, targets.to(cuda_device), tw.to(cuda_device), 1.0, None)

    p = vd.Predictions(images.to(cuda_device), ts, velocities.to(cuda_device))
    nxt, loss = vd.guided_step(p, loss_fn, to_ts, guidance_scale=scale, clamp_value=clampv)

    vr = velocities.clone().requires_grad_()
    loss_ref = guidance_oracle.guidance_loss(dz.denoised_images(images, ts, vr), rows, sd, shape.image_size, shape.patch,
                                             shape.layers, shape.heads, targets, tw, 1.0)
    (gr,) = torch.autograd.grad(loss_ref, vr)
    nxt_ref = dz.step(images, ts, dz.guided(images, ts, velocities, -gr, scale, clampv), to_ts)
    assert abs(float(loss) - float(loss_ref)) <= 1e-2 * abs(float(loss_ref))
    d, dr = (nxt.cpu() - dz.step(images, ts, velocities, to_ts)).double().flatten(), \
        (nxt_ref - dz.step(images, ts, velocities, to_ts)).double().flatten()
    assert float(d @ dr / (d.norm() * dr.norm())) >= 0.999  # the guidance displacement, bf16 path vs fp32 oracle


def test_clamp_with_grad_dtype_and_nan(cuda_device):
    """half inputs keep their dtype (the reference's torch ops do) and NaN propagates like torch.clamp."""
    from perceptor_b200 import transforms
    x = torch.tensor([-0.5, 0.25, 1.5, float("nan")], device=cuda_device)
    y = transforms.clamp_with_grad(x, 0.0, 1.0)
    assert torch.equal(y[:3], torch.tensor([0.0, 0.25, 1.0], device=cuda_device)) and torch.isnan(y[3])
    for dt in (torch.float16, torch.bfloat16):
        xh = torch.tensor([-0.5, 0.25, 1.5], device=cuda_device, dtype=dt).requires_grad_()
        yh = transforms.clamp_with_grad(xh, 0.0, 1.0)
        assert yh.dtype == dt and torch.equal(yh.detach().float(), torch.tensor([0.0, 0.25, 1.0], device=cuda_device))
        yh.sum().backward()
        assert xh.grad.dtype == dt
        # reference rule (clamp_with_grad.py:19-27): grad * [grad * (x - clamp(x)) >= 0] with grad = +1
        assert torch.equal(xh.grad.float(), torch.tensor([0.0, 1.0, 1.0], device=cuda_device))
    p = vd.Predictions(torch.rand(2, 3, 8, 8, device=cuda_device).half(), torch.tensor([0.3, 0.6]),
                       torch.randn(2, 3, 8, 8, device=cuda_device).half())
    assert p.denoised_images.dtype == torch.float16
