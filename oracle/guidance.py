"""Oracle for the whole hot path: cutouts -> resize -> normalize -> ViT -> L2 normalise -> loss, fp32 on CPU.

Composition order follows perceptor/models/open_clip.py:109-123 (resize -> normalize -> encode_image ->
F.normalize) and perceptor/losses/clip/clip.py:89-99.  A cutout is `images[b:b+1, :, y0:y0+s, x0:x0+s]` pushed
through exactly that path.  Normalisation constants: perceptor/models/ruclip/processor.py:23-24.
"""
import torch
import torch.nn.functional as F

from . import loss as loss_oracle
from . import resize as resize_oracle
from . import vit as vit_oracle

MEAN = (0.48145466, 0.4578275, 0.40821073)
STD = (0.26862954, 0.26130258, 0.27577711)


def normalize(x):
    mean = torch.tensor(MEAN, dtype=x.dtype).reshape(1, 3, 1, 1)
    std = torch.tensor(STD, dtype=x.dtype).reshape(1, 3, 1, 1)
    return (x - mean) / std


def cutout_pixels(images, rows, image_size):
    """rows: (b, y0, x0, size) or (b, y0, x0, h, w).  -> [n,3,R,R] normalised."""
    outs = []
    for r in rows:
        b, y0, x0 = int(r[0]), int(r[1]), int(r[2])
        h, w = (int(r[3]), int(r[3])) if len(r) == 4 else (int(r[3]), int(r[4]))
        crop = images[b:b + 1, :, y0:y0 + h, x0:x0 + w]
        outs.append(resize_oracle.resize(crop, (image_size, image_size)))
    return normalize(torch.cat(outs, dim=0))


def encode_cutouts(images, rows, sd, image_size, patch, layers, heads, act="quickgelu", normalize_out=True):
    enc = vit_oracle.encode(cutout_pixels(images, rows, image_size), sd, patch, layers, heads, act)
    return F.normalize(enc) if normalize_out else enc


def guidance_loss(images, rows, sd, image_size, patch, layers, heads, targets, target_weights, multiplier=1.0,
                  act="quickgelu"):
    enc = encode_cutouts(images, rows, sd, image_size, patch, layers, heads, act)
    return loss_oracle.clip_loss(enc, targets, target_weights, multiplier)
