"""S0: the random-cutout sampler (host side).

The reference has no cutout sampler and draws no random numbers on this path (SURVEY.md §0 finding 1): it resizes
every whole image (perceptor/models/open_clip.py:110-118).  "n cutouts" therefore means: n square crops
`images[b:b+1, :, y0:y0+s, x0:x0+s]`, each resized like the reference resizes a whole image.  This module OWNS the
sampler spec; oracle/sampler.py restates it independently and the two must agree bit-for-bit.

Spec (cutouts of image b occupy rows b*n .. b*n+n-1), with N = B*n and one torch.Generator (CPU, mt19937):
    u_size = torch.rand(N, generator=g)   # float32
    u_x    = torch.rand(N, generator=g)
    u_y    = torch.rand(N, generator=g)
    size = int(float64(u_size)**cut_pow * (max_size - min_size) + min_size)       (truncation)
    x0   = int(float64(u_x) * (W - size + 1)),   y0 = int(float64(u_y) * (H - size + 1))
Rows are int32 (b, y0, x0, size).  The three vector draws keep the host cost at three RNG calls per step.
"""
from __future__ import annotations

import numpy as np
import torch


def sample_cutouts(generator: torch.Generator, batch: int, height: int, width: int, n_per_image: int,
                   cut_pow: float = 1.0, min_size: int | None = None, max_size: int | None = None) -> np.ndarray:
    """int32 [batch*n_per_image, 4] rows (b, y0, x0, size)."""
    side = min(height, width)
    max_size = side if max_size is None else int(max_size)
    min_size = min(side, 32) if min_size is None else int(min_size)
    if not (1 <= min_size <= max_size <= side):
        raise ValueError(f"need 1 <= min_size <= max_size <= min(H, W); got {min_size}, {max_size}, {side}")
    if n_per_image <= 0 or batch <= 0:
        raise ValueError("batch and n_per_image must be positive")
    n = batch * n_per_image
    u_size = torch.rand(n, generator=generator).double().numpy()
    u_x = torch.rand(n, generator=generator).double().numpy()
    u_y = torch.rand(n, generator=generator).double().numpy()
    size = (u_size**float(cut_pow) * (max_size - min_size) + min_size).astype(np.int64)
    x0 = (u_x * (width - size + 1)).astype(np.int64)
    y0 = (u_y * (height - size + 1)).astype(np.int64)
    b = np.repeat(np.arange(batch, dtype=np.int64), n_per_image)
    return np.stack([b, y0, x0, size], axis=1).astype(np.int32)


def whole_image_cutouts(batch: int, height: int, width: int) -> np.ndarray:
    """int32 [batch, 5] rows (b, 0, 0, H, W): exactly the reference behaviour (resize every whole image)."""
    rows = np.zeros((batch, 5), dtype=np.int32)
    rows[:, 0] = np.arange(batch)
    rows[:, 3] = height
    rows[:, 4] = width
    return rows


def validate_rows(rows: np.ndarray, batch: int, height: int, width: int) -> np.ndarray:
    """Check caller-supplied cutout rows against the image batch they index; returns them as contiguous int32.

    The sampler kernels address `images + ((b*3 + c)*H + y0 + dy)*W + x0 + dx` straight from the device table and the
    backward scatter-adds into the same addresses, so a row outside the batch would read and WRITE out of bounds.
    Rows are [n,4] (b, y0, x0, size) or [n,5] (b, y0, x0, h, w)."""
    rows = np.asarray(rows)
    if rows.ndim != 2 or rows.shape[1] not in (4, 5) or not np.issubdtype(rows.dtype, np.integer):
        raise ValueError(f"cutout rows must be an integer array of shape [n,4] or [n,5], got {rows.dtype} {rows.shape}")
    r = rows.astype(np.int64, copy=False)
    if r.shape[0]:
        b, y0, x0 = r[:, 0], r[:, 1], r[:, 2]
        h = r[:, 3]
        w = r[:, 3] if r.shape[1] == 4 else r[:, 4]
        bad = (b < 0) | (b >= batch) | (y0 < 0) | (x0 < 0) | (h < 1) | (w < 1) | (y0 + h > height) | (x0 + w > width)
        if bad.any():
            i = int(np.flatnonzero(bad)[0])
            raise ValueError(f"cutout row {i} = {rows[i].tolist()} does not fit a batch of {batch} images of "
                             f"{height}x{width} (need 0 <= b < B, y0, x0 >= 0, sizes >= 1, y0 + h <= H, x0 + w <= W)")
    return np.ascontiguousarray(rows, dtype=np.int32)


def shard_rows(n_rows: int, rank: int, world: int) -> slice:
    """Contiguous block of cutout rows owned by `rank` (§8e: every rank builds the same table, then slices it)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(n_rows, world)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def local_rows(rows: np.ndarray, rank: int, world: int, b_offset: int = 0) -> np.ndarray:
    """The shard of `rank` with the image index re-based by `b_offset` (image-sharded mode: the rank holds only the
    images [b_offset, b_offset + B_local) of the global batch the table was drawn for)."""
    local = rows[shard_rows(rows.shape[0], rank, world)]
    if b_offset:
        local = local.copy()
        local[:, 0] -= b_offset
    return local
