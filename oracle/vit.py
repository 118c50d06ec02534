"""Oracle V1-V4: OpenAI-CLIP VisionTransformer forward, float32 functional torch (differentiable).

Restates perceptor/models/ruclip/model.py: LayerNorm :11-17 (eps 1e-5), QuickGELU :20-22,
ResidualAttentionBlock :25-54 (nn.MultiheadAttention, no mask, need_weights=False), Transformer :57-69,
VisionTransformer.forward :105-131.  Weights come as an OpenAI-style state dict (keys without `visual.`).
"""
import math

import torch
import torch.nn.functional as F


def quick_gelu(x):
    return x * torch.sigmoid(1.702 * x)


def block(x, sd, p, heads, act):
    """x [N,T,D] (batch-first; the reference's LND layout is a pure permutation of the same math)."""
    n, t, d = x.shape
    dh = d // heads
    y = F.layer_norm(x, (d,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], 1e-5)
    qkv = y @ sd[p + "attn.in_proj_weight"].t() + sd[p + "attn.in_proj_bias"]
    q, k, v = (z.reshape(n, t, heads, dh).transpose(1, 2) for z in qkv.chunk(3, dim=-1))
    att = torch.softmax((q / math.sqrt(dh)) @ k.transpose(-1, -2), dim=-1) @ v
    att = att.transpose(1, 2).reshape(n, t, d)
    x = x + att @ sd[p + "attn.out_proj.weight"].t() + sd[p + "attn.out_proj.bias"]
    y = F.layer_norm(x, (d,), sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], 1e-5)
    h = y @ sd[p + "mlp.c_fc.weight"].t() + sd[p + "mlp.c_fc.bias"]
    a = quick_gelu(h) if act == "quickgelu" else F.gelu(h)
    return x + a @ sd[p + "mlp.c_proj.weight"].t() + sd[p + "mlp.c_proj.bias"]


def encode(x, sd, patch, layers, heads, act="quickgelu"):
    """x [N,3,R,R] normalised pixels -> [N,E] (un-normalised encodings)."""
    d = sd["conv1.weight"].shape[0]
    x = F.conv2d(x, sd["conv1.weight"], stride=patch)  # :106
    x = x.reshape(x.shape[0], d, -1).permute(0, 2, 1)  # :107-108
    cls = sd["class_embedding"].to(x.dtype) + torch.zeros(x.shape[0], 1, d, dtype=x.dtype)
    x = torch.cat([cls, x], dim=1) + sd["positional_embedding"]  # :109-119
    x = F.layer_norm(x, (d,), sd["ln_pre.weight"], sd["ln_pre.bias"], 1e-5)  # :120
    for i in range(layers):
        x = block(x, sd, f"transformer.resblocks.{i}.", heads, act)
    x = F.layer_norm(x[:, 0, :], (d,), sd["ln_post.weight"], sd["ln_post.bias"], 1e-5)  # :126
    return x @ sd["proj"]  # :128-129
