"""gradient_checkpoint (perceptor/utils/gradient_checkpoint.py:5-68): evaluate several losses on a detached copy of
a tensor, then continue the backward pass through the common part of the graph once.  Host-side autograd plumbing
only; a plain class stands in for lantern.FunctionalBase."""
from __future__ import annotations

import torch


class GradientCheckpoint:
    def __init__(self, tensor):
        self.original = tensor
        self.detached = tensor.detach().requires_grad_()

    def zero_grad_(self):
        self.detached.grad.zero_()
        return self

    def backward(self, loss):
        loss.backward()
        gradients = self.detached.grad.clone()
        self.zero_grad_()
        return gradients

    def continue_backward(self, gradients=None, retain_graph=False):
        if self.detached.grad is None:
            raise ValueError("Gradient is not defined")
        if gradients is None:
            return self.original.backward(self.detached.grad, retain_graph=retain_graph)
        return self.original.backward(gradients, retain_graph=retain_graph)

    def tensor(self):
        return self.detached

    @staticmethod
    def nonzero_mean(gradients, dim=0):
        if isinstance(gradients, list):
            gradients = torch.stack(gradients)
        return gradients.sum(dim).div(gradients.ne(0).sum(dim).add(1e-6))

    @staticmethod
    def nonzero_scale(tensor, dim=None):
        if isinstance(tensor, list):
            tensor = torch.stack(tensor)
        shape = tensor.shape
        if dim is None:
            tensor = tensor.flatten()
            dim = 0
        mask = tensor.ne(0)
        mean_square = tensor.square().sum(dim) / mask.sum(dim).add(1e-6)
        mean = tensor.sum(dim) / mask.sum(dim).add(1e-6)
        std = (mean_square - mean.square()).sqrt().add(1e-6)
        scaled_tensor = tensor / std.unsqueeze(dim).add(1e-6)
        return scaled_tensor.view(*shape)


def gradient_checkpoint(tensor) -> GradientCheckpoint:
    """
    >>> checkpoint = gradient_checkpoint(images)
    >>> for text_loss in text_losses:
    >>>     text_loss(checkpoint.tensor()).backward()
    >>> checkpoint.continue_backward()
    """
    return GradientCheckpoint(tensor)
