"""CPU oracle for the guided-sampling glue (TEST INFRASTRUCTURE): the reference's eager expressions, written out
one for one in float32 torch.  Pinned by tests/golden/diffusion_glue.npz, which oracle/make_golden_diffusion.py
produces by importing the reference's own predictions.py / velocity_diffusion.py / clamp_with_grad.py.
"""
import math

import torch


def encode(images):  # perceptor/models/velocity_diffusion/diffusion_space.py:1-2
    return images.mul(2).sub(1)


def decode(x):  # diffusion_space.py:5-6
    return x.add(1).div(2)


def t_to_alpha_sigma(t):  # perceptor/models/velocity_diffusion/utils.py:47-50
    return torch.cos(t * math.pi / 2), torch.sin(t * math.pi / 2)


def _planes(v):
    return v[:, None, None, None]


def denoised_xs(images, ts, velocities):  # predictions.py:50-55
    alphas, sigmas = t_to_alpha_sigma(ts)
    return encode(images) * _planes(alphas) - velocities * _planes(sigmas)


def predicted_noise(images, ts, velocities):  # predictions.py:57-62
    alphas, sigmas = t_to_alpha_sigma(ts)
    return encode(images) * _planes(sigmas) + velocities * _planes(alphas)


def denoised_images(images, ts, velocities):  # predictions.py:64-66
    return decode(denoised_xs(images, ts, velocities))


def step(images, ts, velocities, to_ts):  # predictions.py:68-105, eta == 0 branch
    to_alphas, to_sigmas = t_to_alpha_sigma(to_ts)
    to_xs = denoised_xs(images, ts, velocities) * _planes(to_alphas) + predicted_noise(images, ts, velocities) * _planes(
        to_sigmas)
    return decode(to_xs)


def guided(images, ts, velocities, guiding, guidance_scale=0.5, clamp_value=1e-6):  # predictions.py:148-155
    _, sigmas = t_to_alpha_sigma(ts)
    return velocities + guidance_scale * _planes(sigmas) * guiding.clamp(-clamp_value, clamp_value) / clamp_value


def forced_denoised_images(images, ts, velocities, forced):  # predictions.py:176-186 (sigmas >= 1e-3 branch)
    alphas, sigmas = t_to_alpha_sigma(ts)
    den_xs = encode(forced)
    noise = (encode(images) - den_xs * _planes(alphas)) / _planes(sigmas)
    return _planes(alphas) * noise - _planes(sigmas) * den_xs


class _ClampWithGrad(torch.autograd.Function):  # perceptor/transforms/clamp_with_grad.py:8-23
    @staticmethod
    def forward(ctx, input, min, max):
        ctx.min, ctx.max = min, max
        ctx.save_for_backward(input)
        return input.clamp(min, max)

    @staticmethod
    def backward(ctx, grad_in):
        (input,) = ctx.saved_tensors
        return grad_in * (grad_in * (input - input.clamp(ctx.min, ctx.max)) >= 0), None, None


def clamp_with_grad(tensor, min=0.0, max=1.0):  # clamp_with_grad.py:26-27
    return _ClampWithGrad.apply(tensor, min, max)


def dynamic_threshold_velocities(images, ts, velocities, quantile=0.95):  # predictions.py:157-171, N == 1
    den = denoised_xs(images, ts, velocities)
    thr = torch.quantile(den.flatten(start_dim=1).abs(), quantile, dim=1).clamp(min=1.0)
    den = clamp_with_grad(den, -float(thr[0]), float(thr[0])) / thr
    return forced_denoised_images(images, ts, velocities, decode(den))


def schedule_ts(n_steps=500, from_ts=1.0, to_ts=1e-2, rho=7.0):  # velocity_diffusion.py:49-67
    def log_snr(alpha, sigma):  # utils.py:41-44
        return torch.log(alpha**2 / sigma**2)

    fa, fs = t_to_alpha_sigma(torch.as_tensor(from_ts))
    ta, tsig = t_to_alpha_sigma(torch.as_tensor(to_ts))
    hi = (1 / log_snr(fa, fs).exp()).sqrt().clamp(max=150)
    lo = (1 / log_snr(ta, tsig).exp()).sqrt().clamp(min=1e-3)
    ramp = torch.linspace(0, 1, n_steps + 1)
    sigmas = (hi ** (1 / rho) + ramp * (lo ** (1 / rho) - hi ** (1 / rho))) ** rho
    ls = log_snr(torch.ones_like(sigmas), sigmas)
    alpha, sigma = ls.sigmoid().sqrt(), ls.neg().sigmoid().sqrt()  # utils.py:34-38
    t = torch.atan2(sigma, alpha) / math.pi * 2  # utils.py:53-56
    return torch.stack([t[:-1], t[1:]], dim=1)
