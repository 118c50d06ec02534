"""ViT shapes, weight initialisation and weight packing for the native path.

Shapes are the standard OpenAI CLIP vision configs that open_clip builds for the names the reference accepts
(perceptor/losses/clip/clip.py:14-23, perceptor/models/clip.py:6-27); the architecture is the in-tree restatement
perceptor/models/ruclip/model.py:72-131.  State dicts use the OpenAI / open_clip `visual.*` key names (without the
`visual.` prefix) so real checkpoints can be ingested when they are available; offline, weights are random-init
with the same distributions the reference's modules use (ruclip/model.py:93-103 + torch defaults).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import torch

from . import native

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)  # perceptor/models/ruclip/processor.py:23
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)  # perceptor/models/ruclip/processor.py:24


@dataclass(frozen=True)
class VitShape:
    image_size: int
    patch: int
    width: int
    layers: int
    heads: int
    embed: int
    mlp_width: int = 0  # 0: 4 * width (every OpenAI tower); open_clip's ViT-g-14 uses 6144 = 4.36 x 1408

    @property
    def head_dim(self) -> int:
        return self.width // self.heads

    @property
    def head_stride(self) -> int:
        """Columns one head occupies in the qkv / attention-output buffers (include/pcg.h: pcg_head_stride): heads wider
        than 64 (ViT-H/14: 80, ViT-g/14: 88) are zero padded to 128."""
        return 64 if self.head_dim == 64 else 128

    @property
    def grid(self) -> int:
        return self.image_size // self.patch

    @property
    def tokens(self) -> int:
        return self.grid * self.grid + 1

    @property
    def mlp(self) -> int:
        return self.mlp_width or 4 * self.width

    @property
    def kpatch(self) -> int:
        return 3 * self.patch * self.patch

    @property
    def kpad(self) -> int:
        return (self.kpatch + 63) // 64 * 64

    def flops_per_cutout(self, pooled_last_block: bool = False) -> float:
        """Algorithmic forward+backward FLOPs per cutout (dgrad-only backward; SURVEY.md §8d / BASELINE.md §4).

        pooled_last_block=False counts what the reference's dense computation does (every token through every block).
        True counts what the default sequencer executes: in the last block only the K / V projections (and their
        dgrad GEMM) see all T tokens; q, the attention row, the out-projection and the MLP run on the class token.
        """
        g2, d, t, l = self.grid**2, self.width, self.tokens, self.layers
        patch = 2 * g2 * self.kpatch * d
        per_token = 8 * d * d + 4 * d * self.mlp  # QKV 6D^2 + out 2D^2 + MLP 2 * 2 * D * mlp per token
        head = 2 * d * self.embed
        full = l if not pooled_last_block else l - 1
        linear_f = linear_b = full * t * per_token
        attn_f, attn_b = full * 4 * t * t * d, full * 8 * t * t * d
        if pooled_last_block:
            one = 2 * d * d + 2 * d * d + 4 * d * self.mlp          # q, out-projection, MLP of the class-token row
            linear_f += t * 4 * d * d + one                          # K and V of every token
            linear_b += t * 6 * d * d + (2 * d * d + 4 * d * self.mlp)  # full dqkv dgrad GEMM; dout, dfc, dproj on one row
            attn_f += 4 * t * d
            attn_b += 8 * t * d
        return (patch + linear_f + attn_f + head) + (patch + linear_b + attn_b + head)


SHAPES = {
    "ViT-B-32": VitShape(224, 32, 768, 12, 12, 512),
    "ViT-B-16": VitShape(224, 16, 768, 12, 12, 512),
    "ViT-L-14": VitShape(224, 14, 1024, 24, 16, 768),
    "ViT-L-14-336": VitShape(336, 14, 1024, 24, 16, 768),
    # open_clip model_configs/ViT-H-14.json / ViT-g-14.json (head dim 80 / 88: perceptor/models/open_clip.py:24-27,
    # the reference's OpenCLIP default is ViT-H-14, perceptor/losses/open_clip.py:8-12)
    "ViT-H-14": VitShape(224, 14, 1280, 32, 16, 1024),
    "ViT-g-14": VitShape(224, 14, 1408, 40, 16, 1024, mlp_width=6144),
}


def resolve_shape(architecture: str) -> tuple[str, VitShape]:
    """Accepts the reference's spellings (`-quickgelu` suffix, `-336px`)."""
    name = architecture.replace("-quickgelu", "").replace("-336px", "-336")
    if name not in SHAPES:
        raise ValueError(
            f"Invalid architecture: {architecture} (the native path covers {sorted(SHAPES)})")
    return name, SHAPES[name]


def random_state_dict(shape: VitShape, seed: int = 0, dtype=torch.float32) -> dict[str, torch.Tensor]:
    """Random-init fp32 weights under OpenAI key names, same distributions as the reference's modules."""
    g = torch.Generator().manual_seed(seed)
    d, t = shape.width, shape.tokens

    def uniform(shape_, bound):
        return (torch.rand(shape_, generator=g, dtype=dtype) * 2 - 1) * bound

    def normal(shape_, std):
        return torch.randn(shape_, generator=g, dtype=dtype) * std

    sd = {
        "conv1.weight": uniform((d, 3, shape.patch, shape.patch), 1 / math.sqrt(shape.kpatch)),
        "class_embedding": normal((d,), d**-0.5),
        "positional_embedding": normal((t, d), d**-0.5),
        "ln_pre.weight": torch.ones(d, dtype=dtype), "ln_pre.bias": torch.zeros(d, dtype=dtype),
        "ln_post.weight": torch.ones(d, dtype=dtype), "ln_post.bias": torch.zeros(d, dtype=dtype),
        "proj": normal((d, shape.embed), d**-0.5),
    }
    for i in range(shape.layers):
        p = f"transformer.resblocks.{i}."
        sd[p + "ln_1.weight"] = torch.ones(d, dtype=dtype)
        sd[p + "ln_1.bias"] = torch.zeros(d, dtype=dtype)
        sd[p + "ln_2.weight"] = torch.ones(d, dtype=dtype)
        sd[p + "ln_2.bias"] = torch.zeros(d, dtype=dtype)
        sd[p + "attn.in_proj_weight"] = uniform((3 * d, d), math.sqrt(6 / (d + 3 * d)))  # xavier_uniform
        sd[p + "attn.in_proj_bias"] = torch.zeros(3 * d, dtype=dtype)
        sd[p + "attn.out_proj.weight"] = uniform((d, d), 1 / math.sqrt(d))
        sd[p + "attn.out_proj.bias"] = torch.zeros(d, dtype=dtype)
        m = shape.mlp
        sd[p + "mlp.c_fc.weight"] = uniform((m, d), 1 / math.sqrt(d))
        sd[p + "mlp.c_fc.bias"] = uniform((m,), 1 / math.sqrt(d))
        sd[p + "mlp.c_proj.weight"] = uniform((d, m), 1 / math.sqrt(m))
        sd[p + "mlp.c_proj.bias"] = uniform((d,), 1 / math.sqrt(m))
    return sd


def required_keys(layers: int) -> list[str]:
    keys = ["conv1.weight", "class_embedding", "positional_embedding", "ln_pre.weight", "ln_pre.bias",
            "ln_post.weight", "ln_post.bias", "proj"]
    for i in range(layers):
        p = f"transformer.resblocks.{i}."
        keys += [p + s for s in (
            "ln_1.weight", "ln_1.bias", "ln_2.weight", "ln_2.bias", "attn.in_proj_weight", "attn.in_proj_bias",
            "attn.out_proj.weight", "attn.out_proj.bias", "mlp.c_fc.weight", "mlp.c_fc.bias", "mlp.c_proj.weight",
            "mlp.c_proj.bias")]
    return keys


class PackedWeights:
    """Device-resident weights in the layout the kernels consume (include/pcg.h: pcg_vit_weights)."""

    def __init__(self, shape: VitShape, state_dict: dict[str, torch.Tensor], device, act: int):
        self.shape = shape
        self.device = torch.device(device)
        d, kp, kpad = shape.width, shape.kpatch, shape.kpad
        dev, bf = self.device, torch.bfloat16
        self._keep: list[torch.Tensor] = []

        def f32(t):
            t = t.detach().to(dev, torch.float32).contiguous()
            self._keep.append(t)
            return t

        def b16(t):
            t = t.detach().to(dev, torch.float32).to(bf).contiguous()
            self._keep.append(t)
            return t

        def both(w):  # [out,in] and its transpose, bf16
            w = w.detach().to(dev, torch.float32)
            return b16(w), b16(w.t())

        sd = state_dict
        conv = sd["conv1.weight"].detach().to(dev, torch.float32).reshape(d, kp)
        conv_pad = torch.zeros(d, kpad, device=dev)
        conv_pad[:, :kp] = conv
        self.conv1, self.conv1_t = both(conv_pad)
        self.cls = f32(sd["class_embedding"])
        self.pos = f32(sd["positional_embedding"])
        self.ln_pre_g, self.ln_pre_b = f32(sd["ln_pre.weight"]), f32(sd["ln_pre.bias"])
        self.ln_post_g, self.ln_post_b = f32(sd["ln_post.weight"]), f32(sd["ln_post.bias"])
        self.proj = f32(sd["proj"])

        self.layers_c = (native.LayerWeights * shape.layers)()
        for i in range(shape.layers):
            p = f"transformer.resblocks.{i}."
            w_in = sd[p + "attn.in_proj_weight"].detach().to(dev, torch.float32).clone()
            b_in = sd[p + "attn.in_proj_bias"].detach().to(dev, torch.float32).clone()
            # fold the 1/sqrt(head_dim) attention scale into the q projection (1/8 for head dim 64: exact in bf16)
            w_in[:d] *= shape.head_dim ** -0.5
            b_in[:d] *= shape.head_dim ** -0.5
            w_o = sd[p + "attn.out_proj.weight"].detach().to(dev, torch.float32)
            if shape.head_stride != shape.head_dim:
                # wide heads: head h of q / k / v moves to rows [h*128, h*128 + head_dim) of its third, the rest are zero
                # rows (so the pad columns of qkv are exactly zero); out_proj gets matching zero columns
                hd, hs, nh = shape.head_dim, shape.head_stride, shape.heads
                w_pad = torch.zeros(3, nh, hs, d, device=dev)
                w_pad[:, :, :hd] = w_in.reshape(3, nh, hd, d)
                b_pad = torch.zeros(3, nh, hs, device=dev)
                b_pad[:, :, :hd] = b_in.reshape(3, nh, hd)
                o_pad = torch.zeros(d, nh, hs, device=dev)
                o_pad[:, :, :hd] = w_o.reshape(d, nh, hd)
                w_in, b_in, w_o = w_pad.reshape(3 * nh * hs, d), b_pad.reshape(3 * nh * hs), o_pad.reshape(d, nh * hs)
            w_qkv, w_qkv_t = both(w_in)
            w_out, w_out_t = both(w_o)
            w_fc, w_fc_t = both(sd[p + "mlp.c_fc.weight"])
            w_proj, w_proj_t = both(sd[p + "mlp.c_proj.weight"])
            lw = self.layers_c[i]
            for name, t in (
                ("ln1_g", f32(sd[p + "ln_1.weight"])), ("ln1_b", f32(sd[p + "ln_1.bias"])),
                ("ln2_g", f32(sd[p + "ln_2.weight"])), ("ln2_b", f32(sd[p + "ln_2.bias"])),
                ("w_qkv", w_qkv), ("w_qkv_t", w_qkv_t), ("w_out", w_out), ("w_out_t", w_out_t),
                ("w_fc", w_fc), ("w_fc_t", w_fc_t), ("w_proj", w_proj), ("w_proj_t", w_proj_t),
                ("b_qkv", f32(b_in)), ("b_out", f32(sd[p + "attn.out_proj.bias"])),
                ("b_fc", f32(sd[p + "mlp.c_fc.bias"])), ("b_proj", f32(sd[p + "mlp.c_proj.bias"])),
            ):
                setattr(lw, name, t.data_ptr())

        self.cfg = native.VitConfig(
            image_size=shape.image_size, patch=shape.patch, grid=shape.grid, tokens=shape.tokens, width=shape.width,
            layers=shape.layers, heads=shape.heads, mlp=shape.mlp, embed=shape.embed, kpatch=shape.kpatch,
            kpad=shape.kpad, act=act, head_dim=shape.head_dim)
        self.weights_c = native.VitWeights(
            conv1=self.conv1.data_ptr(), conv1_t=self.conv1_t.data_ptr(), cls=self.cls.data_ptr(),
            pos=self.pos.data_ptr(), ln_pre_g=self.ln_pre_g.data_ptr(), ln_pre_b=self.ln_pre_b.data_ptr(),
            ln_post_g=self.ln_post_g.data_ptr(), ln_post_b=self.ln_post_b.data_ptr(), proj=self.proj.data_ptr(),
            layers_host=C.cast(self.layers_c, C.POINTER(native.LayerWeights)))

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self._keep)
