"""Encoder wrappers with the reference's surface: models.CLIP(...) / models.OpenCLIP(...).

Mirrors perceptor/models/clip.py:6-27 and perceptor/models/open_clip.py:11-137 for the vision side of the hot
path: `.encode_images(images, normalize=True)`, `.image_size`, `.device`, frozen eval-mode weights.  The image
encoder runs on the native sm_100a path (bf16 tensor-core GEMMs, fp32 residual stream / statistics); offline there
are no pretrained weights, so weights are random-init unless a `state_dict` is supplied.
"""
from __future__ import annotations

import dataclasses
import warnings
import weakref

import numpy as np
import torch
import torch.nn.functional as F

from . import checkpoints, cutouts, native, text
from .guidance import EncodeImagesFn, GuidanceEngine
from .vit import random_state_dict, required_keys, resolve_shape

# (architecture, weights) pairs the reference documents for ViT towers
# (perceptor/models/open_clip.py:24-44); anything else raises ValueError like the reference does (:52-53).
_PRETRAINED = {
    ("ViT-B-32-quickgelu", "openai"), ("ViT-B-32", "openai"), ("ViT-B-16", "openai"), ("ViT-L-14", "openai"),
    ("ViT-L-14-336", "openai"),
    ("ViT-B-32", "laion2b_s34b_b79k"), ("ViT-B-32", "laion2b_e16"), ("ViT-B-32", "laion400m_e32"),
    ("ViT-B-16", "laion400m_e32"), ("ViT-L-14", "laion2b_s32b_b82k"), ("ViT-L-14", "laion400m_e32"),
    # head dim 80 / 88: heads padded to 128 columns, attention on the mma.sync kernels (perceptor/models/open_clip.py:24-27)
    ("ViT-H-14", "laion2b_s32b_b79k"), ("ViT-g-14", "laion2b_s12b_b42k"),
}


_precision_warned: set[str] = set()


class _OpenCLIP(torch.nn.Module):
    def __init__(self, architecture="ViT-L-14", weights="openai", precision=None, *, state_dict=None, seed=0,
                 bpe_path=None):
        super().__init__()
        self._bpe_path = bpe_path
        self._seed = int(seed)
        self._tokenizer = None
        self._text_sd = None
        self.architecture = architecture
        self.weights = weights
        if (architecture, weights) not in _PRETRAINED:
            raise ValueError(f"Invalid architecture/weights: {architecture}/{weights}")
        if precision not in (None, "fp32", "fp16", "bf16"):
            raise ValueError(f"Invalid precision: {precision}")
        # The reference honours `precision` (perceptor/models/open_clip.py:56-63: None -> fp16 on CUDA, "fp32" kept).
        # The native path has ONE arithmetic: bf16 tensor-core operands, fp32 accumulation / residual stream /
        # LayerNorm and softmax statistics / loss.  Asking for anything else is served by that path and says so once
        # (north_star's parity bars -- loss 1e-2 relative, gradient cosine 0.999 against the fp32 path -- hold).
        self.requested_precision = precision
        self.precision = "bf16"
        if precision in ("fp32", "fp16") and precision not in _precision_warned:
            _precision_warned.add(precision)
            warnings.warn(f"perceptor_b200: precision={precision!r} requested; the native sm_100a path computes with bf16 "
                          "tensor-core operands and fp32 accumulation / statistics (no fp32 or fp16 kernels exist). "
                          "`.precision` reports 'bf16', `.requested_precision` keeps the argument.", stacklevel=3)
        self.name, self.shape = resolve_shape(architecture)
        quick = weights == "openai" or "-quickgelu" in architecture
        self.act = native.ACT_QUICKGELU if quick else native.ACT_GELU
        # OpenAI / open_clip / Hugging Face layouts -> OpenAI vision names (+ "text.*" when the text tower is there)
        sd = checkpoints.normalize_state_dict(state_dict) if state_dict is not None else random_state_dict(self.shape, seed)
        self._text_shape = text.TEXT_SHAPES[self.name]
        if all("text." + k in sd for k in text.text_keys(self._text_shape.layers)):
            self._text_sd = {k: sd["text." + k].detach().float() for k in text.text_keys(self._text_shape.layers)}
        missing = [k for k in required_keys(self.shape.layers) if k not in sd]
        if missing:
            raise ValueError(f"state_dict is missing vision keys: {missing[:4]}...")
        self._keys = required_keys(self.shape.layers)
        for k in self._keys:
            self.register_parameter(k.replace(".", "__"), torch.nn.Parameter(sd[k].detach().float(), requires_grad=False))
        self.eval()
        self._engines: dict[torch.device, GuidanceEngine] = {}
        start_device = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
        if start_device.type == "cuda":
            self.to(start_device)

    # -- reference surface ------------------------------------------------------------------------------
    @property
    def device(self):
        return next(iter(self.parameters())).device

    @property
    def image_size(self):
        return (self.shape.image_size, self.shape.image_size)

    def state_dict_openai(self) -> dict[str, torch.Tensor]:
        return {k: getattr(self, k.replace(".", "__")).detach() for k in self._keys}

    @torch.no_grad()
    def encode_texts(self, text_prompts, normalize=True):
        """perceptor/models/open_clip.py:99-107: tokenize -> text transformer -> (L2-normalise).  Plain PyTorch on
        the encoder's device (once per prompt).  Needs the CLIP BPE merge table: `bpe_path=` at
        construction or PCG_BPE_VOCAB.  Without a checkpoint the text tower is random-init like the image tower."""
        if self._tokenizer is None:
            self._tokenizer = text.SimpleTokenizer(self._bpe_path)
        if self._text_sd is None:
            self._text_sd = text.random_text_state_dict(self._text_shape, self._seed + 1)
        dev = self.device
        if next(iter(self._text_sd.values())).device != dev:
            self._text_sd = {k: v.to(dev) for k, v in self._text_sd.items()}
        tokens = text.tokenize(self._tokenizer, text_prompts, self._text_shape.context)
        encodings = text.encode_text(self._text_sd, self._text_shape, tokens, quick_gelu=self.act == native.ACT_QUICKGELU)
        return F.normalize(encodings) if normalize else encodings

    def engine(self) -> GuidanceEngine:
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("the native CLIP image path needs a CUDA device; there is no CPU fallback")
        dev = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        eng = self._engines.get(dev)
        if eng is None:
            eng = GuidanceEngine(self.shape, self.state_dict_openai(), dev, self.act)
            self._engines = {dev: eng}  # one device at a time; drop stale packs
        return eng

    def encode_images(self, images, normalize=True, *, cutout_rows=None):
        """[N,3,H,W] -> [N,E].  Each whole image is resized to image_size (the reference behaviour); with
        `cutout_rows` ([n,4] int rows b,y0,x0,size) one encoding per cutout is returned instead."""
        eng = self.engine()
        images = images.to(eng.device)
        if images.dtype != torch.float32:
            images = images.float()
        images = images.contiguous()
        if cutout_rows is None:
            cutout_rows = cutouts.whole_image_cutouts(images.shape[0], images.shape[2], images.shape[3])
        else:
            cutout_rows = cutouts.validate_rows(cutout_rows, images.shape[0], images.shape[2], images.shape[3])
        plan = eng.plan_cutouts(np.asarray(cutout_rows))
        return EncodeImagesFn.apply(images, eng, plan, bool(normalize))

    @staticmethod
    def spherical_distance(encodings_a, encodings_b):
        return (encodings_a[:, None] - encodings_b[None, :]).norm(dim=2).div(2).arcsin().square().mul(2)

    def forward(self, _):
        raise NotImplementedError


_cache: "weakref.WeakValueDictionary[str, _OpenCLIP]" = weakref.WeakValueDictionary()


def OpenCLIP(architecture="ViT-L-14", weights="openai", precision=None, *, state_dict=None, seed=0, bpe_path=None):
    """Weak-valued memoised constructor (perceptor/utils/cache.py:9-23): equal arguments share one encoder."""
    if state_dict is not None:
        return _OpenCLIP(architecture, weights, precision, state_dict=state_dict, bpe_path=bpe_path)
    key = str((architecture, weights, precision, int(seed), bpe_path))
    model = _cache.get(key)
    if model is None:
        model = _OpenCLIP(architecture, weights, precision, seed=seed, bpe_path=bpe_path)
        _cache[key] = model
    return model


def CLIP(architecture: str, precision=None, **kwargs):
    """perceptor/models/clip.py:6-27: OpenAI weights; B/32 (and RN50/RN101, not covered) get `-quickgelu`."""
    if "-quickgelu" not in architecture and architecture in ["RN50", "RN101", "ViT-B-32"]:
        architecture = architecture + "-quickgelu"
    return OpenCLIP(architecture, "openai", precision, **kwargs)


def normalize_encodings(encodings: torch.Tensor) -> torch.Tensor:
    return F.normalize(encodings)


# ---------------------------------------------------------------------------------------------------------
# TransformersOpenAICLIP surface (perceptor/models/transformers_openai_clip.py:17-137): the same
# encoders addressed by Hugging Face model id, `encode_images` / `encode_texts` returning an `Encodings` record.
# ---------------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class Encodings:
    """perceptor/models/transformers_openai_clip.py:17-21.  `features` (HF's BaseModelOutputWithPooling with every hidden
    state) is not produced by the native path: the engine keeps activations in its own stash layout, so it is None."""
    features: object
    unnormalized_encodings: torch.Tensor
    encodings: torch.Tensor


# Hugging Face id -> (architecture, weights) of the towers this package builds (perceptor/models/
# transformers_openai_clip.py:36-50 lists more: the XLM-Roberta / sentence-transformers text towers are out of scope)
_HF_IDS = {
    "openai/clip-vit-base-patch32": ("ViT-B-32-quickgelu", "openai"),
    "openai/clip-vit-base-patch16": ("ViT-B-16", "openai"),
    "openai/clip-vit-large-patch14": ("ViT-L-14", "openai"),
    "openai/clip-vit-large-patch14-336": ("ViT-L-14-336", "openai"),
    "laion/CLIP-ViT-L-14-laion2B-s32B-b82K": ("ViT-L-14", "laion2b_s32b_b82k"),
    "laion/CLIP-ViT-B-32-laion2B-s34B-b79K": ("ViT-B-32", "laion2b_s34b_b79k"),
    "laion/CLIP-ViT-H-14-laion2B-s32B-b79K": ("ViT-H-14", "laion2b_s32b_b79k"),
    "laion/CLIP-ViT-g-14-laion2B-s12B-b42K": ("ViT-g-14", "laion2b_s12b_b42k"),
}


class TransformersOpenAICLIP(torch.nn.Module):
    """`TransformersOpenAICLIP(name)` of the reference with the native image tower behind it.  There is no network here:
    pass the checkpoint as `state_dict=` (Hugging Face, open_clip or OpenAI layout; `checkpoints.normalize_state_dict`)
    or get random-init weights.  `bfloat16` is accepted for signature parity; the native path is bf16 either way."""

    def __init__(self, name="openai/clip-vit-large-patch14", bfloat16=True, *, state_dict=None, seed=0, bpe_path=None):
        super().__init__()
        if name not in _HF_IDS:
            raise ValueError(f"Invalid model id: {name} (known: {sorted(_HF_IDS)})")
        self.name = name
        architecture, weights = _HF_IDS[name]
        self.clip = OpenCLIP(architecture, weights, "bf16", state_dict=state_dict, seed=seed, bpe_path=bpe_path)
        # (the reference takes 224 from the base-patch32 feature extractor for every id, :71-75; the native tower's
        # positional embedding fixes the size, so the 336 model resizes to 336 here)
        self.image_size = list(self.clip.image_size)

    @property
    def device(self):
        return self.clip.device

    def encode_images(self, images) -> Encodings:
        unnormalized = self.clip.encode_images(images, normalize=False)
        return Encodings(None, unnormalized, unnormalized / unnormalized.norm(p=2, dim=-1, keepdim=True))

    def encode_texts(self, texts) -> Encodings:
        unnormalized = self.clip.encode_texts(texts, normalize=False)
        return Encodings(None, unnormalized, unnormalized / unnormalized.norm(p=2, dim=-1, keepdim=True))

    @staticmethod
    def spherical_distance(encodings_a: Encodings, encodings_b: Encodings) -> torch.Tensor:
        return (encodings_a.encodings[:, None] - encodings_b.encodings[None, :]).norm(dim=2).div(2).arcsin().square().mul(2)

    def forward(self, _):
        raise NotImplementedError
