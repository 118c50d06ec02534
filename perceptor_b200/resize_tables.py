"""S1 host side: per-size tap tables for the separable antialiased resize.

The tables (first input index + weights per output pixel) are computed with the SAME sequence of float32 torch
operations the reference uses, so the integer indices are bit-identical to
perceptor/transforms/resize/resize_right.py (get_projected_grid :192-207, apply_antialiasing_if_needed :426-436,
get_field_of_view :210-219, calc_pad_sz :222-233, get_weights :275-285; kernels interpolation_methods.py:38-62).
The CUDA sampler only applies them.  Default reference kwargs are assumed (resample=None, antialiasing=True,
by_convs=False, pad_mode="constant"), which is how the hot path calls resize (perceptor/models/open_clip.py:113-116).
"""
from __future__ import annotations

from dataclasses import dataclass
from math import ceil, pi

import numpy as np
import torch

LANCZOS3, CUBIC = 0, 1
_SUPPORT = {LANCZOS3: 6, CUBIC: 4}
_EPS = float(torch.finfo(torch.float32).eps)


def _cubic(x: torch.Tensor) -> torch.Tensor:
    absx = torch.abs(x)
    absx2 = absx**2
    absx3 = absx**3
    return (1.5 * absx3 - 2.5 * absx2 + 1.0) * (absx <= 1.0).to(x.dtype) + (
        -0.5 * absx3 + 2.5 * absx2 - 4.0 * absx + 2.0
    ) * ((1.0 < absx) & (absx <= 2.0)).to(x.dtype)


def _lanczos3(x: torch.Tensor) -> torch.Tensor:
    return ((torch.sin(pi * x) * torch.sin(pi * x / 3) + _EPS) / ((pi**2 * x**2 / 3) + _EPS)) * (abs(x) < 3).to(x.dtype)


_KERNELS = {LANCZOS3: _lanczos3, CUBIC: _cubic}


def choose_method(in_h: int, in_w: int, out_h: int, out_w: int) -> int:
    """resize_right.py:102-108: lanczos3 when neither dim is upscaled, bicubic otherwise."""
    return LANCZOS3 if (in_h >= out_h and in_w >= out_w) else CUBIC


@dataclass(frozen=True)
class DimTable:
    """One resized dimension: output o reads inputs left[o] .. left[o]+taps-1 (indices outside [0,in_size) are the
    zero padding) with weights[o, :]."""

    in_size: int
    out_size: int
    method: int
    taps: int
    left: np.ndarray     # int32 [out]
    weights: np.ndarray  # float32 [out, taps]
    inv: np.ndarray      # int32 [in, 2]: outputs [lo, hi) that read input i


def build_dim_table(in_size: int, out_size: int, method: int) -> DimTable:
    if in_size <= 0 or out_size <= 0:
        raise ValueError(f"resize sizes must be positive, got {in_size} -> {out_size}")
    scale = out_size / in_size  # resize_right.py:376-378 (out_shape given, scale derived)
    if float(scale) == 1.0:
        # the reference skips dims whose scale is exactly 1.0 (resize_right.py:114-118)
        left = np.arange(out_size, dtype=np.int32)
        weights = np.ones((out_size, 1), dtype=np.float32)
        taps = 1
    else:
        support = _SUPPORT[method]
        kernel = _KERNELS[method]
        out_coordinates = torch.arange(out_size)
        projected_grid = out_coordinates / float(scale) + (in_size - 1) / 2 - (out_size - 1) / (2 * float(scale))
        if scale >= 1.0:
            cur_kernel, cur_support = kernel, support
        else:  # antialiasing: stretch the kernel by 1/scale
            cur_kernel = lambda arg: scale * kernel(scale * arg)  # noqa: E731
            cur_support = support / scale
        left_boundaries = (projected_grid - cur_support / 2 - _EPS).ceil().long()
        taps = ceil(cur_support - _EPS)
        field_of_view = left_boundaries[:, None] + torch.arange(taps)
        # the reference pads the input so that the smallest index becomes 0 and shifts both coordinate systems by
        # the same integer (calc_pad_sz); replicate the shift so the float32 rounding of the weights is identical.
        pad0 = -int(field_of_view[0, 0].item())
        shifted_fov = field_of_view + pad0
        shifted_grid = projected_grid + pad0
        w = cur_kernel(shifted_grid[:, None] - shifted_fov)
        sum_w = w.sum(1, keepdim=True)
        sum_w[sum_w == 0] = 1
        w = w / sum_w
        left = left_boundaries.to(torch.int32).numpy()
        weights = w.to(torch.float32).contiguous().numpy()
    idx = np.arange(in_size)
    lo = np.searchsorted(left + taps - 1, idx, side="left")
    hi = np.searchsorted(left, idx, side="right")
    inv = np.stack([lo, np.maximum(hi, lo)], axis=1).astype(np.int32)
    return DimTable(in_size, out_size, method, taps, left.astype(np.int32), weights.astype(np.float32), inv)


class ResizeTableCache:
    """Host cache + device mirror of DimTables for one output size.  ids index `desc` on the device."""

    def __init__(self, out_size: int):
        self.out_size = int(out_size)
        self._ids: dict[tuple[int, int], int] = {}
        self._tables: list[DimTable] = []
        self._device = None
        self._dev_tensors = None
        self._dirty = True

    def table_id(self, in_size: int, method: int) -> int:
        key = (int(in_size), int(method))
        tid = self._ids.get(key)
        if tid is None:
            tid = len(self._tables)
            self._tables.append(build_dim_table(key[0], self.out_size, key[1]))
            self._ids[key] = tid
            self._dirty = True
        return tid

    def table(self, tid: int) -> DimTable:
        return self._tables[tid]

    def ensure_sizes(self, sizes, method_for_size) -> None:
        for s in sizes:
            self.table_id(int(s), method_for_size(int(s)))

    def flat_arrays(self):
        """(desc int32 [n,8], left int32, weight float32, inv int32) in the layout include/pcg.h documents."""
        desc = np.zeros((len(self._tables), 8), dtype=np.int32)
        lefts, weights, invs = [], [], []
        lo = wo = io = 0
        for i, t in enumerate(self._tables):
            desc[i, :5] = (t.taps, lo, wo, io, t.in_size)
            lefts.append(t.left)
            weights.append(t.weights.reshape(-1))
            invs.append(t.inv.reshape(-1))
            lo += t.left.size
            wo += t.weights.size
            io += t.inv.size
        return (desc, np.concatenate(lefts) if lefts else np.zeros(0, np.int32),
                np.concatenate(weights) if weights else np.zeros(0, np.float32),
                np.concatenate(invs) if invs else np.zeros(0, np.int32))

    def device_tensors(self, device):
        """Upload (again) when tables were added; returns (desc, left, weight, inv) CUDA tensors."""
        if self._dirty or self._device != device:
            desc, left, weight, inv = self.flat_arrays()
            self._dev_tensors = tuple(
                torch.from_numpy(np.ascontiguousarray(a)).to(device) for a in (desc, left, weight, inv))
            self._device = device
            self._dirty = False
        return self._dev_tensors
