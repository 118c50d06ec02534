"""Oracle S1: separable antialiased resize, float32 torch on CPU, differentiable.

Restates perceptor/transforms/resize/resize_right.py for the kwargs the hot path uses (resample=None,
antialiasing=True, by_convs=False, pad_mode="constant"; perceptor/models/open_clip.py:113-116):
resize() :34-189, get_projected_grid :192-207, get_field_of_view :210-219, calc_pad_sz :222-233,
get_weights :275-285, apply_weights :288-318, apply_antialiasing_if_needed :426-436, and the lanczos3 / cubic
kernels of interpolation_methods.py:38-46,57-62.
"""
from math import ceil, pi

import torch
import torch.nn.functional as F

EPS = float(torch.finfo(torch.float32).eps)


def cubic(x):  # interpolation_methods.py:38-46
    a = x.abs()
    return (1.5 * a**3 - 2.5 * a**2 + 1.0) * (a <= 1.0).to(x.dtype) + (
        -0.5 * a**3 + 2.5 * a**2 - 4.0 * a + 2.0) * ((1.0 < a) & (a <= 2.0)).to(x.dtype)


def lanczos3(x):  # interpolation_methods.py:57-62
    return ((torch.sin(pi * x) * torch.sin(pi * x / 3) + EPS) / ((pi**2 * x**2 / 3) + EPS)) * (abs(x) < 3).to(x.dtype)


KERNELS = {"lanczos3": (lanczos3, 6), "cubic": (cubic, 4)}


def dim_taps(in_size, out_size, method):
    """(field_of_view int64 [out,taps] in UNPADDED input coordinates, weights f32 [out,taps]) for one dim."""
    kernel, support = KERNELS[method]
    scale = float(out_size / in_size)
    grid = torch.arange(out_size) / scale + (in_size - 1) / 2 - (out_size - 1) / (2 * scale)  # :203-207
    if scale < 1.0:  # :426-436
        base, s = kernel, scale
        kernel = lambda a: s * base(s * a)  # noqa: E731
        support = support / scale
    left = (grid - support / 2 - EPS).ceil().long()  # :214
    fov = left[:, None] + torch.arange(ceil(support - EPS))  # :218-219
    pad0 = -int(fov[0, 0])  # :226-231: both coordinate systems move by the left pad
    w = kernel((grid + pad0)[:, None] - (fov + pad0))  # :279
    s_w = w.sum(1, keepdim=True)
    s_w[s_w == 0] = 1
    return fov, w / s_w  # :282-285


def resize_dim(x, dim, out_size, method):
    in_size = x.shape[dim]
    if float(out_size / in_size) == 1.0:  # :114-118 (dims with scale 1.0 are skipped)
        return x
    fov, w = dim_taps(in_size, out_size, method)
    pad_l, pad_r = -int(fov[0, 0]), int(fov[-1, -1]) - in_size + 1  # may be negative = crop (:226)
    xt = x.transpose(dim, -1)
    xt = F.pad(xt, (pad_l, pad_r), mode="constant", value=0.0).transpose(dim, -1)  # fw_pad :466-480
    xt = xt.transpose(dim, 0)
    neighbors = xt[fov + pad_l]  # [out, taps, ...]  :305
    w = w.to(x.dtype).reshape(*w.shape, *([1] * (x.dim() - 1)))
    return (neighbors * w).sum(1).transpose(0, dim)  # :314-318


def resize(images, out_shape):
    """images [N,C,H,W] -> [N,C,out_h,out_w]."""
    h, w = images.shape[-2:]
    oh, ow = out_shape
    method = "lanczos3" if (h >= oh and w >= ow) else "cubic"  # :102-108
    dims = sorted([(oh / h, 2, oh), (ow / w, 3, ow)], key=lambda t: t[0])  # ascending scale, stable (:114-118)
    out = images
    for _, dim, size in dims:
        out = resize_dim(out, dim, size, method)
    return out
