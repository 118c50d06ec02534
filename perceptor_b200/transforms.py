"""clamp_with_grad (perceptor/transforms/clamp_with_grad.py:8-41): clamp in the forward pass; in the backward
pass a gradient is only blocked when it would push an already-clamped value further out of range.  Forward and
backward are one native elementwise kernel each (csrc/diffusion.cu); CUDA tensors only."""
from __future__ import annotations

import torch

from . import native


def _launch(x, g, lo, hi):
    if not x.is_cuda:
        raise RuntimeError("the native clamp_with_grad needs CUDA tensors; there is no CPU fallback")
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):  # launch on the tensor's GPU and ITS current stream
        native.check(native.lib().pcg_clamp_with_grad(native.ptr(x), native.ptr(g), native.ptr(out), x.numel(), float(lo),
                                                      float(hi), native.stream_ptr()), "pcg_clamp_with_grad")
    return out


class ClampWithGradFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, min=0, max=1):
        ctx.min, ctx.max = float(min), float(max)
        ctx.dtype = input.dtype  # fp32 kernel; half / bf16 inputs get their dtype back (the reference preserves it)
        input = input.contiguous().float()
        ctx.save_for_backward(input)
        return _launch(input, None, ctx.min, ctx.max).to(ctx.dtype)

    @staticmethod
    def backward(ctx, grad_in):
        (input,) = ctx.saved_tensors
        return _launch(input, grad_in.contiguous().float(), ctx.min, ctx.max).to(ctx.dtype), None, None


def clamp_with_grad(tensor, min=0.0, max=1.0):
    return ClampWithGradFunction.apply(tensor, min, max)


class ClampWithGrad(torch.nn.Module):
    """TransformInterface-shaped wrapper (encode clamps, decode is the identity)."""

    def __init__(self, min=0, max=1):
        super().__init__()
        self.min = min
        self.max = max

    def encode(self, tensor):
        return clamp_with_grad(tensor, self.min, self.max)

    def decode(self, tensor):
        return tensor

    def forward(self, tensor):
        return self.encode(tensor)
