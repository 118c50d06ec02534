"""perceptor_b200 — B200-native CLIP-guidance loss (perceptor.losses.CLIP / OpenCLIP hot path).

    from perceptor_b200 import losses
    loss = losses.CLIP("ViT-L-14", n_cutouts=128).add_encodings_(text_encodings)
    loss(images).backward()          # images.grad: [B,3,H,W]

Beside the hot path (losses, models, cutouts, resize_tables, vit, guidance, native) the package carries the callers'
glue it was widened into: velocity_diffusion (Predictions, schedule_ts, guided_step), transforms
(clamp_with_grad), text (CLIP tokenizer + text tower) and checkpoints (OpenAI /
open_clip / Hugging Face state-dict ingestion).
"""
from . import checkpoints, cutouts, losses, models, resize_tables, text, transforms, velocity_diffusion, vit  # noqa: F401

__all__ = ["losses", "models", "cutouts", "resize_tables", "vit", "velocity_diffusion", "transforms", "text",
           "checkpoints"]
