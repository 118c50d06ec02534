"""Oracle S0: independent restatement of the cutout-sampler spec (perceptor_b200/cutouts.py docstring).

The reference has no sampler (SURVEY.md §0 finding 1); parity here is oracle-vs-product only.  Scalar Python
arithmetic (doubles) on the same three torch.Generator draws, deliberately not vectorised.
"""
import torch


def sample_cutouts(generator, batch, height, width, n_per_image, cut_pow=1.0, min_size=None, max_size=None):
    side = min(height, width)
    max_size = side if max_size is None else int(max_size)
    min_size = min(side, 32) if min_size is None else int(min_size)
    n = batch * n_per_image
    u_size = torch.rand(n, generator=generator).tolist()
    u_x = torch.rand(n, generator=generator).tolist()
    u_y = torch.rand(n, generator=generator).tolist()
    rows = []
    for i in range(n):
        size = int(u_size[i] ** float(cut_pow) * (max_size - min_size) + min_size)
        x0 = int(u_x[i] * (width - size + 1))
        y0 = int(u_y[i] * (height - size + 1))
        rows.append((i // n_per_image, y0, x0, size))
    return rows
