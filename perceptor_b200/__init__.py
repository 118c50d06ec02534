"""perceptor_b200 — B200-native CLIP-guidance loss (perceptor.losses.CLIP / OpenCLIP hot path).

    from perceptor_b200 import losses
    loss = losses.CLIP("ViT-L-14", n_cutouts=128).add_encodings_(text_encodings)
    loss(images).backward()          # images.grad: [B,3,H,W]
"""
from . import cutouts, losses, models, resize_tables, vit  # noqa: F401

__all__ = ["losses", "models", "cutouts", "resize_tables", "vit"]
