"""Guided-sampling glue either side of the guidance loss, with the reference's names.

Mirrors, without `lantern`:
  * perceptor/models/velocity_diffusion/diffusion_space.py:1-6      encode / decode
  * perceptor/models/velocity_diffusion/utils.py:28-78              alpha/sigma/log-snr/t conversions
  * perceptor/models/velocity_diffusion/velocity_diffusion.py:49-67 schedule_ts (Karras ramp in elucidated sigma)
  * perceptor/models/velocity_diffusion/predictions.py:9-197        Predictions
The per-sample affine maps (`denoised_images`, `step(eta=0)`, `guided`) run as ONE fused native kernel each
(csrc/diffusion.cu) with an autograd.Function around it, so that `clip_loss(predictions.denoised_images)`
back-propagates into `velocities` / `from_diffused_images` exactly like the reference's eager chain.  There is no
CPU fallback: CPU tensors raise, like every other native entry point.  The UNet that produces `velocities` is the
caller (config 5) and is not part of this package.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, replace

import torch

from . import native
from .transforms import clamp_with_grad


# ---------------------------------------------------------------------------------------------------------
# diffusion_space.py / utils.py
# ---------------------------------------------------------------------------------------------------------
def encode(images):
    return images.mul(2).sub(1)


def decode(x):
    return x.add(1).div(2)


def log_snr_to_alpha_sigma(log_snr):
    return log_snr.sigmoid().sqrt(), log_snr.neg().sigmoid().sqrt()


def alpha_sigma_to_log_snr(alpha, sigma):
    return torch.log(alpha**2 / sigma**2)


def t_to_alpha_sigma(t):
    return torch.cos(t * math.pi / 2), torch.sin(t * math.pi / 2)


def alpha_sigma_to_t(alpha, sigma):
    return torch.atan2(sigma, alpha) / math.pi * 2


def sigma_to_t(sigma):
    return torch.asin(sigma) / math.pi * 2


def schedule_ts(n_steps=500, from_ts=1.0, to_ts=1e-2, rho=7.0) -> torch.Tensor:
    """[n_steps, 2] rows (from_t, to_t): perceptor/models/velocity_diffusion/velocity_diffusion.py:49-67."""
    from_alpha, from_sigma = t_to_alpha_sigma(torch.as_tensor(from_ts))
    to_alpha, to_sigma = t_to_alpha_sigma(torch.as_tensor(to_ts))
    from_log_snr = alpha_sigma_to_log_snr(from_alpha, from_sigma)
    to_log_snr = alpha_sigma_to_log_snr(to_alpha, to_sigma)
    elucidated_from_sigma = (1 / from_log_snr.exp()).sqrt().clamp(max=150)
    elucidated_to_sigma = (1 / to_log_snr.exp()).sqrt().clamp(min=1e-3)
    ramp = torch.linspace(0, 1, n_steps + 1)
    min_inv_rho = elucidated_to_sigma ** (1 / rho)
    max_inv_rho = elucidated_from_sigma ** (1 / rho)
    sigmas = (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho
    log_snr = alpha_sigma_to_log_snr(torch.ones_like(sigmas), sigmas)
    alpha, sigma = log_snr_to_alpha_sigma(log_snr)
    ts = alpha_sigma_to_t(alpha, sigma)
    return torch.stack([ts[:-1], ts[1:]], dim=1)


# ---------------------------------------------------------------------------------------------------------
# the fused affine map and its autograd
# ---------------------------------------------------------------------------------------------------------
def _coef(device, *rows) -> torch.Tensor:
    """[3, N] float32 coefficient table on `device` (rows computed wherever the timesteps live)."""
    coef = torch.stack([r.reshape(-1).float() for r in rows])
    if coef.device != device:
        coef = coef.pin_memory().to(device, non_blocking=True) if coef.device.type == "cpu" else coef.to(device)
    return coef.contiguous()


def _launch_affine(x, y, coef, limit):
    if not x.is_cuda:
        raise RuntimeError("the native guided-sampling path needs CUDA tensors; there is no CPU fallback")
    dtype = x.dtype  # the kernel computes in fp32; the result goes back to the caller's dtype like the reference's torch ops
    x = x.contiguous().float()
    y = None if y is None else y.contiguous().float()
    out = torch.empty_like(x)
    n = coef.shape[1]
    if x.shape[0] != n:
        raise ValueError(f"batch {x.shape[0]} does not match {n} timesteps")
    with torch.cuda.device(x.device):  # launch on the tensor's GPU and ITS current stream, not the process default
        native.check(native.lib().pcg_affine2(native.ptr(x), native.ptr(y), native.ptr(coef), native.ptr(out), n,
                                              x.numel() // n, float(limit), native.stream_ptr()), "pcg_affine2")
    return out if dtype == torch.float32 or not dtype.is_floating_point else out.to(dtype)


class _Affine2(torch.autograd.Function):
    """out = a[n] x + b[n] clamp(y, -limit, limit) + c[n]; gradients flow to x and (when limit is inf) to y."""

    @staticmethod
    def forward(ctx, x, y, coef, limit):
        ctx.save_for_backward(coef)
        ctx.has_y = y is not None
        ctx.limit = limit
        return _launch_affine(x, y, coef, limit)

    @staticmethod
    def backward(ctx, grad):
        (coef,) = ctx.saved_tensors
        zeros = torch.zeros_like(coef[0])
        gx = gy = None
        if ctx.needs_input_grad[0]:
            gx = _launch_affine(grad, None, torch.stack([coef[0], zeros, zeros]), math.inf)
        if ctx.has_y and ctx.needs_input_grad[1]:
            if math.isfinite(ctx.limit):
                raise RuntimeError("no gradient flows through the clamped guiding term (the reference clamps it too)")
            gy = _launch_affine(grad, None, torch.stack([coef[1], zeros, zeros]), math.inf)
        return gx, gy, None, None


def affine2(x, y, a, b, c, limit=math.inf):
    return _Affine2.apply(x, y, _coef(x.device, a, b, c), limit)


def _ts_1d(ts) -> torch.Tensor:
    """predictions.py:18-25: a float or 0-d tensor becomes a 1-element vector; anything but 1-D is an error."""
    if isinstance(ts, float):
        ts = torch.tensor(ts)
    if ts.ndim == 0:
        ts = ts[None]
    if ts.ndim != 1:
        raise ValueError("ts must be a scalar or a 1D tensor")
    return ts


@dataclass(frozen=True)
class Predictions:
    """perceptor/models/velocity_diffusion/predictions.py:9-197 (a frozen dataclass stands in for lantern.FunctionalBase)."""

    from_diffused_images: torch.Tensor  # NCHW in [0, 1]
    from_ts: torch.Tensor               # N
    velocities: torch.Tensor            # NCHW

    def replace(self, **kwargs) -> "Predictions":
        return replace(self, **kwargs)

    @property
    def device(self):
        return self.velocities.device

    def alphas(self, ts) -> torch.Tensor:
        alphas, _ = t_to_alpha_sigma(_ts_1d(ts))
        return alphas[:, None, None, None].to(self.device)

    def sigmas(self, ts) -> torch.Tensor:
        _, sigmas = t_to_alpha_sigma(_ts_1d(ts))
        return sigmas[:, None, None, None].to(self.device)

    @property
    def from_alphas(self):
        return self.alphas(self.from_ts)

    @property
    def from_sigmas(self):
        return self.sigmas(self.from_ts)

    @property
    def from_diffused_xs(self):
        return encode(self.from_diffused_images)

    def _n(self, ts) -> torch.Tensor:
        """timesteps broadcast to the batch (a single timestep serves every sample)"""
        ts = _ts_1d(ts)
        n = self.velocities.shape[0]
        return ts.expand(n) if ts.shape[0] == 1 and n != 1 else ts

    @property
    def denoised_xs(self):
        """X alpha - v sigma with X = 2 x - 1 (predictions.py:50-55)."""
        alpha, sigma = t_to_alpha_sigma(self._n(self.from_ts))
        return affine2(self.from_diffused_images, self.velocities, 2 * alpha, -sigma, -alpha)

    @property
    def predicted_noise(self):
        alpha, sigma = t_to_alpha_sigma(self._n(self.from_ts))
        # eps = X sigma + v alpha with X = 2 x - 1
        return affine2(self.from_diffused_images, self.velocities, 2 * sigma, alpha, -sigma)

    @property
    def denoised_images(self):
        """decode(X alpha - v sigma) as one fused kernel (predictions.py:50-66)."""
        alpha, sigma = t_to_alpha_sigma(self._n(self.from_ts))
        return affine2(self.from_diffused_images, self.velocities, alpha, -sigma / 2, (1 - alpha) / 2)

    def step(self, to_ts, eta=0.0):
        """Reduce the noise level to `to_ts` (predictions.py:68-105); eta = 0 is one fused kernel."""
        alpha, sigma = t_to_alpha_sigma(self._n(self.from_ts))
        to_alpha, to_sigma = t_to_alpha_sigma(self._n(to_ts).to(alpha.device))
        if eta > 0.0:
            ddim_sigma = eta * (to_sigma**2 / sigma**2).sqrt() * (1 - alpha**2 / to_alpha**2).sqrt()
            noise_scale = (to_sigma**2 - ddim_sigma**2).sqrt()
        else:
            ddim_sigma, noise_scale = None, to_sigma
        big_a = alpha * to_alpha + sigma * noise_scale
        out = affine2(self.from_diffused_images, self.velocities, big_a, (alpha * noise_scale - sigma * to_alpha) / 2,
                      (1 - big_a) / 2)
        if ddim_sigma is not None:
            out = out + torch.randn_like(out) * (ddim_sigma[:, None, None, None].to(out.device) / 2)
        return out

    def correction(self, previous: "Predictions") -> "Predictions":
        return previous.forced_denoised_images((self.denoised_images + previous.denoised_images) / 2)

    def reverse_step(self, to_ts):
        if (torch.as_tensor(self.from_ts) > torch.as_tensor(to_ts)).any():
            raise ValueError("from_ts must be less than to_ts")
        return self.denoised_xs * self.alphas(to_ts) + self.predicted_noise * self.sigmas(to_ts)

    def resample_noise(self, resample_ts):
        if (torch.as_tensor(self.from_ts) < torch.as_tensor(resample_ts)).any():
            raise ValueError("from_ts must be greater than resample_ts")
        eps = self.predicted_noise
        resampled = self.sigmas(resample_ts) * eps + (
            self.from_sigmas**2 - self.sigmas(resample_ts) ** 2).sqrt() * torch.randn_like(eps)
        return resampled / self.from_sigmas

    def resample(self, resample_ts):
        return decode(self.denoised_xs * self.from_alphas + self.resample_noise(resample_ts) * self.from_sigmas)

    def noisy_reverse_step(self, to_ts):
        to_alphas, to_sigmas = self.alphas(to_ts), self.sigmas(to_ts)
        eps = self.predicted_noise
        noise_sigma = self.from_sigmas * eps + (to_sigmas**2 - self.from_sigmas**2).sqrt() * torch.randn_like(eps)
        return decode(self.denoised_xs * to_alphas + noise_sigma)

    def guided(self, guiding, guidance_scale=0.5, clamp_value=1e-6) -> "Predictions":
        """v + scale sigma clamp(g, -c, c) / c as one fused kernel (predictions.py:148-155)."""
        _, sigma = t_to_alpha_sigma(self._n(self.from_ts))
        return self.replace(velocities=affine2(self.velocities, guiding.detach(), torch.ones_like(sigma),
                                               guidance_scale * sigma / clamp_value, torch.zeros_like(sigma),
                                               limit=clamp_value))

    def dynamic_threshold(self, quantile=0.95) -> "Predictions":
        """Thresholding heuristic from the imagen paper (predictions.py:157-171).  The reference broadcasts its [N]
        threshold against NCHW, which is only meaningful for N = 1; here every sample gets its own threshold (the
        same numbers for N = 1: clamp(x, -t, t) / t == clamp(x / t, -1, 1) for t > 0, gradients included)."""
        denoised_xs = self.denoised_xs
        threshold = torch.quantile(denoised_xs.flatten(start_dim=1).abs(), quantile, dim=1).clamp(min=1.0)
        threshold = threshold[:, None, None, None]
        clamped = clamp_with_grad(denoised_xs / threshold, -1.0, 1.0)
        return self.forced_denoised_images(decode(clamped))

    def static_threshold(self) -> "Predictions":
        return self.forced_denoised_images(clamp_with_grad(self.denoised_images, 0, 1))

    def forced_denoised_images(self, denoised_images) -> "Predictions":
        denoised_xs = encode(denoised_images)
        if (self.from_sigmas >= 1e-3).all():
            predicted_noise = (self.from_diffused_xs - denoised_xs * self.from_alphas) / self.from_sigmas
        else:
            predicted_noise = self.predicted_noise
        return self.replace(velocities=self.from_alphas * predicted_noise - self.from_sigmas * denoised_xs)

    def forced_predicted_noise(self, predicted_noise) -> "Predictions":
        if (self.from_alphas >= 1e-3).all():
            denoised_xs = (self.from_diffused_xs - predicted_noise * self.from_sigmas) / self.from_alphas
        else:
            denoised_xs = self.denoised_xs
        return self.replace(velocities=self.from_alphas * predicted_noise - self.from_sigmas * denoised_xs)


def diffuse(denoised_images, ts, noise=None):
    """velocity_diffusion.py:138-146: decode(encode(x) alpha + noise sigma)."""
    ts = _ts_1d(ts if not isinstance(ts, float) else torch.tensor(ts))
    n = denoised_images.shape[0]
    ts = ts.expand(n) if ts.shape[0] == 1 and n != 1 else ts
    if noise is None:
        noise = torch.randn_like(denoised_images)
    alpha, sigma = t_to_alpha_sigma(ts)
    return affine2(denoised_images, noise, alpha, sigma / 2, (1 - alpha) / 2)


def guided_step(predictions: Predictions, loss_fn, to_ts, guidance_scale=0.5, clamp_value=1e-6, eta=0.0):
    """One CLIP-guided sampling step (the inner loop of config 5): the loss is evaluated on the denoised images of
    `predictions`, its gradient w.r.t. the velocities guides them (the README's `predictions.guided(-grad)` usage),
    and the guided predictions step to `to_ts`.  Returns (diffused_images at to_ts, loss value)."""
    with torch.enable_grad():
        velocities = predictions.velocities.detach().requires_grad_()
        tracked = predictions.replace(velocities=velocities)
        loss = loss_fn(tracked.denoised_images)
        (grad,) = torch.autograd.grad(loss, velocities)
    guided = predictions.guided(-grad, guidance_scale=guidance_scale, clamp_value=clamp_value)
    return guided.step(to_ts, eta=eta), loss.detach()
