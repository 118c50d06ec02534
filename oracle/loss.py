"""Oracle E1/E2/T1: target handling and the spherical-distance loss.

Restates perceptor/losses/clip/clip.py:60-99 (CLIP: targets re-normalised, multiplier 0.01 for ViT-L names
:31-34) and perceptor/losses/open_clip.py:58-97 (OpenCLIP: targets stored as given, no multiplier).
"""
import torch


def spherical_distance(image_encodings, target_encodings):  # clip.py:91-98
    return (image_encodings[:, None] - target_encodings[None, :]).norm(dim=2).div(2).arcsin().square().mul(2)


def clip_loss(image_encodings, target_encodings, target_weights, multiplier=1.0):  # clip.py:99
    return (spherical_distance(image_encodings, target_encodings) * target_weights).mean().mul(multiplier)


def default_multiplier(name):  # clip.py:31-34
    return 0.01 if name in ("ViT-L-14", "ViT-L-14-336") else 1.0
