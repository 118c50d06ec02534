"""Checkpoint ingestion: OpenAI / open_clip / Hugging Face CLIP state dicts -> the OpenAI key
names the native engine packs (`vit.required_keys`) plus the text-tower keys `text.text_keys` uses.

The reference loads weights through open_clip.create_model_and_transforms (perceptor/models/open_clip.py:65-72) or
transformers.CLIPModel (perceptor/models/transformers_openai_clip.py:53-69); offline there is nothing to download,
so the entry point is a state dict the user already has:

    sd = checkpoints.normalize_state_dict(torch.load("ViT-L-14.pt"))        # any of the three layouts
    loss = losses.CLIP("ViT-L-14", state_dict=sd, bpe_path=...)

Layouts
  * OpenAI / open_clip: `visual.*` for the image tower, text tower at the top level -> split, `visual.` stripped.
  * vision-only OpenAI names (what `vit.random_state_dict` produces) -> returned as is.
  * Hugging Face (`vision_model.*`, `visual_projection.weight`, `text_model.*`, `text_projection.weight`):
    q/k/v projections are concatenated into `attn.in_proj_*`, `pre_layrnorm` (sic) is ln_pre, the projection
    matrices are transposed (HF stores nn.Linear weights, OpenAI stores `x @ proj`).
"""
from __future__ import annotations

import re

import torch

_HF_VISION_FIXED = {
    "vision_model.embeddings.class_embedding": "class_embedding",
    "vision_model.embeddings.patch_embedding.weight": "conv1.weight",
    "vision_model.embeddings.position_embedding.weight": "positional_embedding",
    "vision_model.pre_layrnorm.weight": "ln_pre.weight", "vision_model.pre_layrnorm.bias": "ln_pre.bias",
    "vision_model.post_layernorm.weight": "ln_post.weight", "vision_model.post_layernorm.bias": "ln_post.bias",
}
_HF_TEXT_FIXED = {
    "text_model.embeddings.token_embedding.weight": "token_embedding.weight",
    "text_model.embeddings.position_embedding.weight": "positional_embedding",
    "text_model.final_layer_norm.weight": "ln_final.weight", "text_model.final_layer_norm.bias": "ln_final.bias",
}
_HF_LAYER = {
    "layer_norm1": "ln_1", "layer_norm2": "ln_2", "self_attn.out_proj": "attn.out_proj", "mlp.fc1": "mlp.c_fc",
    "mlp.fc2": "mlp.c_proj",
}


def _hf_tower(sd, prefix: str) -> dict[str, torch.Tensor]:
    """`{prefix}.encoder.layers.N.*` -> `transformer.resblocks.N.*` with fused in_proj."""
    out = {}
    pat = re.compile(rf"^{re.escape(prefix)}\.encoder\.layers\.(\d+)\.(.+)\.(weight|bias)$")
    qkv: dict[tuple[int, str], dict[str, torch.Tensor]] = {}
    for key, val in sd.items():
        m = pat.match(key)
        if not m:
            continue
        layer, name, kind = int(m.group(1)), m.group(2), m.group(3)
        if name in ("self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj"):
            qkv.setdefault((layer, kind), {})[name[-6]] = val
        elif name in _HF_LAYER:
            out[f"transformer.resblocks.{layer}.{_HF_LAYER[name]}.{kind}"] = val
    for (layer, kind), parts in qkv.items():
        out[f"transformer.resblocks.{layer}.attn.in_proj_{kind}"] = torch.cat([parts["q"], parts["k"], parts["v"]])
    return out


def is_hf(sd) -> bool:
    return any(k.startswith("vision_model.") for k in sd)


def normalize_state_dict(sd) -> dict[str, torch.Tensor]:
    """Returns {vision keys under OpenAI names} | {"text.<key>": text-tower tensors when the checkpoint has them}."""
    sd = {k[len("model."):] if k.startswith("model.") else k: v for k, v in sd.items()}
    out: dict[str, torch.Tensor] = {}
    if is_hf(sd):
        for src, dst in _HF_VISION_FIXED.items():
            if src in sd:
                out[dst] = sd[src]
        out.update(_hf_tower(sd, "vision_model"))
        if "visual_projection.weight" in sd:
            out["proj"] = sd["visual_projection.weight"].t().contiguous()
        if "text_model.embeddings.token_embedding.weight" in sd:
            for src, dst in _HF_TEXT_FIXED.items():
                out["text." + dst] = sd[src]
            out.update({"text." + k: v for k, v in _hf_tower(sd, "text_model").items()})
            if "text_projection.weight" in sd:
                out["text.text_projection"] = sd["text_projection.weight"].t().contiguous()
        return out
    if any(k.startswith("visual.") for k in sd):
        for k, v in sd.items():
            if k.startswith("visual."):
                out[k[len("visual."):]] = v
            elif k.startswith(("transformer.", "token_embedding.", "ln_final.")) or k in ("positional_embedding",
                                                                                          "text_projection"):
                out["text." + k] = v
        return out
    return dict(sd)
