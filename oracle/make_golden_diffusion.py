"""Generate tests/golden/diffusion_glue.npz from the REFERENCE's own files (build container only).

perceptor/models/velocity_diffusion/{predictions,velocity_diffusion,utils,diffusion_space}.py and
perceptor/transforms/clamp_with_grad.py are imported UNMODIFIED by path; the packages they import but that are not
installed offline (lantern, basicsr, the `perceptor` package root) are replaced by minimal stand-ins that provide
only the names those files touch (a keyword-constructed record with .replace() for lantern.FunctionalBase).
"""
import importlib
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference/perceptor")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def _pkg(name, path=None):
    m = types.ModuleType(name)
    m.__path__ = [str(path)] if path else []
    sys.modules[name] = m
    return m


def load_reference():
    class FunctionalBase:
        def __init__(self, **kw):
            self.__dict__.update(kw)

        def replace(self, **kw):
            d = dict(self.__dict__)
            d.update(kw)
            return type(self)(**d)

    class _Tensor:
        @staticmethod
        def dims(*_a, **_k):
            return torch.Tensor

        @staticmethod
        def shape(*_a, **_k):
            return torch.Tensor

    lantern = _pkg("lantern")
    lantern.FunctionalBase, lantern.Tensor = FunctionalBase, _Tensor
    _pkg("basicsr"), _pkg("basicsr.utils")
    dl = _pkg("basicsr.utils.download_util")
    dl.load_file_from_url = lambda *a, **k: None
    perceptor = _pkg("perceptor")
    utils = _pkg("perceptor.utils")
    utils.cache = lambda f: f
    perceptor.utils = utils
    perceptor.models = _pkg("perceptor.models")
    _pkg("perceptor.transforms", REF / "transforms")
    cwg = importlib.import_module("perceptor.transforms.clamp_with_grad")
    _pkg("perceptor.models.velocity_diffusion", REF / "models" / "velocity_diffusion")
    stub_models = _pkg("perceptor.models.velocity_diffusion.models")
    stub_models.get_model = lambda name: None
    pred = importlib.import_module("perceptor.models.velocity_diffusion.predictions")
    vd = importlib.import_module("perceptor.models.velocity_diffusion.velocity_diffusion")
    return pred.Predictions, vd.VelocityDiffusion, cwg.clamp_with_grad


def main():
    Predictions, VelocityDiffusion, clamp_with_grad = load_reference()
    g = torch.Generator().manual_seed(11)
    out = {}
    out["schedule_50"] = VelocityDiffusion.schedule_ts(n_steps=50).numpy()
    out["schedule_7"] = VelocityDiffusion.schedule_ts(n_steps=7, from_ts=0.9, to_ts=0.05, rho=5.0).numpy()

    images = torch.rand(3, 3, 10, 12, generator=g)
    velocities = torch.randn(3, 3, 10, 12, generator=g)
    ts = torch.tensor([0.9, 0.5, 0.13])
    to_ts = torch.tensor([0.8, 0.41, 0.02])
    guiding = torch.randn(3, 3, 10, 12, generator=g) * 2e-6
    p = Predictions(from_diffused_images=images, from_ts=ts, velocities=velocities)
    out.update(images=images.numpy(), velocities=velocities.numpy(), ts=ts.numpy(), to_ts=to_ts.numpy(),
               guiding=guiding.numpy(), denoised_xs=p.denoised_xs.numpy(), predicted_noise=p.predicted_noise.numpy(),
               denoised_images=p.denoised_images.numpy(), step=p.step(to_ts).numpy(),
               guided=p.guided(guiding, guidance_scale=0.7, clamp_value=1e-6).velocities.numpy(),
               forced=p.forced_denoised_images(images.flip(0)).velocities.numpy())
    # gradient of a scalar of denoised_images w.r.t. the velocities and the diffused images
    v = velocities.clone().requires_grad_()
    x = images.clone().requires_grad_()
    cot = torch.randn(3, 3, 10, 12, generator=g)
    q = Predictions(from_diffused_images=x, from_ts=ts, velocities=v)
    gv, gx = torch.autograd.grad((q.denoised_images * cot).sum(), (v, x))
    out.update(cot=cot.numpy(), grad_velocities=gv.numpy(), grad_images=gx.numpy())
    # single-sample thresholding (the reference's [N] threshold only broadcasts for N == 1)
    p1 = Predictions(from_diffused_images=images[:1], from_ts=ts[:1], velocities=velocities[:1] * 3)
    out["dynamic_threshold"] = p1.dynamic_threshold(0.9).velocities.numpy()
    out["static_threshold"] = p1.static_threshold().velocities.numpy()
    # clamp_with_grad forward / backward
    z = (torch.randn(4, 50, generator=g) * 0.8 + 0.5).requires_grad_()
    gz_in = torch.randn(4, 50, generator=g)
    y = clamp_with_grad(z, 0.0, 1.0)
    (gz,) = torch.autograd.grad(y, z, gz_in)
    out.update(cwg_x=z.detach().numpy(), cwg_y=y.detach().numpy(), cwg_gin=gz_in.numpy(), cwg_gx=gz.numpy())
    np.savez_compressed(OUT / "diffusion_glue.npz", **out)
    print("diffusion_glue.npz", (OUT / "diffusion_glue.npz").stat().st_size)


if __name__ == "__main__":
    main()
