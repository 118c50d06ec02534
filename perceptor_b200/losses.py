"""losses.CLIP / losses.OpenCLIP with the reference's module API, backed by the native guidance path.

API parity (perceptor/losses/clip/clip.py:10-99, perceptor/losses/open_clip.py:7-97): construct with a model name,
`add_texts_` / `add_images_` / `add_encodings_` / `add_text_off_` / `mul_` return self, `encodings` / `weights` are
frozen nn.Parameters rebuilt by torch.cat, `forward(images[N,3,H,W])` returns a 0-d loss whose autograd gradient
flows into `images`.  Keyword-only extras (n_cutouts, cut_pow, min_size, max_size, seed/generator, process_group,
shard) default to the reference behaviour: every whole image is resized and encoded.

Multi-GPU (`process_group=`): `shard="cutouts"` (default) -- every rank passes the SAME images, the cutout table is
split over the ranks and the image gradient is summed with one all-reduce; `shard="images"` -- every rank passes its
OWN images (same count and size), the loss is the mean over all ranks' cutouts and the gradient needs no collective.
"""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

from . import cutouts, models
from .guidance import GuidanceLossFn


class _TextImageLoss(torch.nn.Module):
    """Shared machinery; subclasses set the reference quirks (re-normalising targets, multiplier)."""

    _renormalize_targets = True

    def _init_guidance(self, n_cutouts, cut_pow, min_size, max_size, seed, generator, process_group, shard="cutouts"):
        if shard not in ("cutouts", "images"):
            raise ValueError(f"shard must be 'cutouts' or 'images', got {shard!r}")
        self.shard = shard
        self.n_cutouts = n_cutouts
        self.cut_pow = float(cut_pow)
        self.min_size = min_size
        self.max_size = max_size
        self.process_group = process_group
        self.generator = generator if generator is not None else torch.Generator().manual_seed(int(seed))
        # the module's own generator may be drawn from one call ahead (see _loss); a caller's generator never is
        self._own_generator = generator is None
        self._ahead = None
        self.encodings = None
        self.weights = None
        self.last_cutouts: np.ndarray | None = None

    @property
    def device(self):
        return next(iter(self.model.parameters())).device

    def add_texts_(self, texts, weights=None):
        return self.add_encodings_(self.model.encode_texts(texts), weights)

    def add_images_(self, images, weights=None):
        with torch.no_grad():
            return self.add_encodings_(self.model.encode_images(images), weights)

    def add_encodings_(self, encodings, weights=None):
        if isinstance(weights, (list, tuple)):
            weights = torch.tensor(weights)
        elif weights is None:
            weights = torch.ones_like(encodings[:, 0])
        encodings = encodings.detach().float()
        if self._renormalize_targets:
            encodings = F.normalize(encodings)
        encodings = encodings.to(self.device)
        weights = weights.detach().float().to(self.device)
        if self.encodings is None:
            self.encodings = torch.nn.Parameter(encodings, requires_grad=False)
            self.weights = torch.nn.Parameter(weights, requires_grad=False)
        else:
            self.encodings = torch.nn.Parameter(torch.cat([self.encodings, encodings]), requires_grad=False)
            self.weights = torch.nn.Parameter(torch.cat([self.weights, weights]), requires_grad=False)
        return self

    # ------------------------------------------------------------------------------------------------
    def _cutout_rows(self, images) -> np.ndarray:
        b, _, h, w = images.shape
        return self._cutout_rows_for(b, h, w)

    def _cutout_rows_for(self, b: int, h: int, w: int) -> np.ndarray:
        if self.n_cutouts is None:
            return cutouts.whole_image_cutouts(b, h, w)
        return cutouts.sample_cutouts(self.generator, b, h, w, int(self.n_cutouts), self.cut_pow, self.min_size,
                                      self.max_size)

    def _loss(self, images, multiplier: float):
        if self.encodings is None:
            raise ValueError("no targets: call add_texts_/add_images_/add_encodings_ first")
        eng = self.model.engine()
        images = images.to(eng.device)
        if images.dtype != torch.float32:
            images = images.float()
        images = images.contiguous()
        group = self.process_group
        rank = world = None
        if group is not None:
            rank, world = torch.distributed.get_rank(group), torch.distributed.get_world_size(group)
        targets = self.encodings.detach().to(eng.device, torch.float32).contiguous()
        tweights = self.weights.detach().to(eng.device, torch.float32).contiguous()
        stream = torch.cuda.current_stream(eng.device).cuda_stream
        if group is not None and world > 1 and self.shard == "images":
            # image-sharded: every rank passes ITS images (the same count and size on every rank).  The table is drawn
            # for the concatenated batch from the common seed, so the result equals the single-process loss over all
            # ranks' images; a rank's cutouts touch only its own images, so the gradient needs no collective.
            b_local = images.shape[0]
            draw = (b_local * world, images.shape[2], images.shape[3])
            b_offset, reduce_grad = rank * b_local, False
        else:
            draw = (images.shape[0], images.shape[2], images.shape[3])
            b_offset, reduce_grad = None, True
        key = (draw, rank or 0, world or 1, b_offset, id(eng), stream)
        rows, plan = self._take_ahead(key)
        if rows is None:
            rows, plan = self._draw_and_plan(eng, draw, rank or 0, world or 1, b_offset)
        self.last_cutouts = rows
        loss = GuidanceLossFn.apply(images, eng, plan, targets, tweights, float(multiplier), group, reduce_grad)
        if self._own_generator and self.n_cutouts is not None:
            state = self.generator.get_state()
            rows, plan = self._draw_and_plan(eng, draw, rank or 0, world or 1, b_offset)
            self._ahead = {"key": key, "state": state, "state_after": self.generator.get_state(), "rows": rows,
                           "plan": plan, "n_cutouts": self.n_cutouts, "spec": self._cutout_spec()}
        return loss

    def _draw_and_plan(self, eng, draw, rank, world, b_offset):
        rows = self._cutout_rows_for(*draw)
        if b_offset is not None and rows.shape[0] % world != 0:
            raise ValueError("image-sharded mode needs the same number of cutouts on every rank")
        return rows, eng.plan_cutouts(rows, rank, world, b_offset=b_offset)

    # The forward above only QUEUES work on the GPU.  Drawing and planning the next call's cutouts right away puts that
    # host work (~0.3 ms) under the GPU time of this call instead of in front of the next one -- which matters to a
    # caller that reads the loss back every step.  The draws are the ones the next call would make anyway; if the next
    # call turns out different (shape, world), the generator is put back to where it was and nothing has changed; if the
    # caller has touched `self.generator` in between (reseeded it, drawn from it), the look-ahead is dropped and the
    # generator is used as found.  (A state saved with `generator.get_state()` after a call is one draw ahead.)
    def _take_ahead(self, key):
        ahead, self._ahead = self._ahead, None
        if ahead is None:
            return None, None
        if not torch.equal(self.generator.get_state(), ahead["state_after"]):
            return None, None  # the caller reseeded (or drew from) the generator since: its state is the one that counts
        if ahead["key"] == key and ahead["n_cutouts"] == self.n_cutouts and ahead["spec"] == self._cutout_spec():
            return ahead["rows"], ahead["plan"]
        self.generator.set_state(ahead["state"])
        return None, None

    def _cutout_spec(self):
        return (self.cut_pow, self.min_size, self.max_size)


class CLIP(_TextImageLoss):
    def __init__(self, name="ViT-B-32", precision="fp32", jit=False, *, n_cutouts=None, cut_pow=1.0, min_size=None,
                 max_size=None, seed=0, generator=None, process_group=None, shard="cutouts", state_dict=None,
                 weights_seed=0, bpe_path=None):
        """
        Args:
            name: name of the clip model. Available models on the native path:
                - ViT-B-32
                - ViT-B-16
                - ViT-L-14
                - ViT-L-14-336
        """
        super().__init__()
        self.name = name
        extra = {"seed": weights_seed} if state_dict is None else {"state_dict": state_dict}
        self.model = models.CLIP(name, precision, bpe_path=bpe_path, **extra)
        self._init_guidance(n_cutouts, cut_pow, min_size, max_size, seed, generator, process_group, shard)
        self.multiplier = 0.01 if name in ("ViT-L-14", "ViT-L-14-336") else 1.0

    def mul_(self, multiplier):
        self.multiplier *= multiplier
        return self

    def add_text_off_(self, weight=None, path="perceptor/losses/clip/vectors/textoff.json"):
        textoff_json = json.loads(Path(path).read_text())
        if self.name in textoff_json:
            return self.add_encodings_(torch.tensor(textoff_json[self.name]), weight)
        raise ValueError(f"There is no textoff for this model: {self.name}")

    def forward(self, images):
        return self._loss(images, self.multiplier)


class OpenCLIP(_TextImageLoss):
    _renormalize_targets = False  # perceptor/losses/open_clip.py:58-85 stores encodings as given

    def __init__(self, architecture="ViT-H-14", weights="laion2b_s32b_b79k", *, n_cutouts=None, cut_pow=1.0,
                 min_size=None, max_size=None, seed=0, generator=None, process_group=None, shard="cutouts",
                 state_dict=None, weights_seed=0, bpe_path=None):
        """
        Args:
            architecture (str): name of the clip model
            weights (str): name of the weights

        Defaults as in the reference (perceptor/losses/open_clip.py:8-12).  ViT-H-14 / ViT-g-14 have head dim 80 / 88:
        their heads are padded to 128 columns and attention runs on the mma.sync kernels (the tcgen05 attention kernels
        are head-dim-64: ViT-B/32, B/16, L/14, L/14@336).
        """
        super().__init__()
        self.architecture = architecture
        extra = {"seed": weights_seed} if state_dict is None else {"state_dict": state_dict}
        self.model = models.OpenCLIP(architecture, weights, bpe_path=bpe_path, **extra)
        self._init_guidance(n_cutouts, cut_pow, min_size, max_size, seed, generator, process_group, shard)

    def forward(self, images):
        return self._loss(images, 1.0)


# ---------------------------------------------------------------------------------------------------------
# other consumers of the same image encoder: the native path ends at `encode_images`, whose
# autograd.Function takes an arbitrary upstream gradient; the heads below are a few dozen FLOPs of torch on [N, E].
# ---------------------------------------------------------------------------------------------------------
class SphericalDistance(torch.nn.Module):
    """perceptor/losses/spherical_distance.py:4-21: mean pairwise 2 asin(|a - b| / 2)^2 between two image batches."""

    def __init__(self, model):
        super().__init__()
        self.model = model

    def forward(self, images_a, images_b):
        enc_a, enc_b = self.model.encode_images(images_a), self.model.encode_images(images_b)
        return (enc_a[:, None] - enc_b[None, :]).norm(dim=2).div(2).arcsin().square().mul(2).mean()


class SimulacraAesthetic(torch.nn.Module):
    """perceptor/losses/simulacra_aesthetic.py:8-44 over perceptor/models/simulacra_aesthetic/simulacra_aesthetic.py:
    26-60: a linear probe on sqrt(E) * normalised CLIP encodings predicts the rating; loss = multiplier * mse to the
    target.  The probe's checkpoint cannot be downloaded offline: pass `head_state_dict` ({"linear.weight",
    "linear.bias"}) or get a random-init probe."""

    def __init__(self, model_name="ViT-L-14", aesthetic_target=10, *, head_state_dict=None, state_dict=None,
                 weights_seed=0):
        super().__init__()
        self.aesthetic_target = torch.nn.Parameter(torch.as_tensor(aesthetic_target).float(), requires_grad=False)
        extra = {"seed": weights_seed} if state_dict is None else {"state_dict": state_dict}
        self.clip_model = models.CLIP(model_name, **extra)
        self.linear = torch.nn.Linear(self.clip_model.shape.embed, 1)
        if head_state_dict is not None:
            self.linear.load_state_dict({k.replace("linear.", ""): v for k, v in head_state_dict.items()})
        self.linear.eval().requires_grad_(False)
        self.linear.to(self.clip_model.device)
        self.multiplier = 0.00001 if model_name in ("ViT-L-14", "ViT-L-14-336") else 0.001

    def predict(self, images):
        encodings = self.clip_model.encode_images(images)
        return self.linear(F.normalize(encodings, dim=-1) * encodings.shape[-1] ** 0.5)

    def forward(self, images):
        return self.multiplier * F.mse_loss(self.predict(images), self.aesthetic_target.view(-1, 1).to(self.linear.weight))


class AestheticVisualAssessment(torch.nn.Module):
    """perceptor/losses/aesthetic_visual_assessment.py:10-51: a 10-way rating classifier on ViT-B/16 encodings."""

    def __init__(self, aesthetic_target=10, mode="expected", *, head_state_dict=None, state_dict=None, weights_seed=0):
        super().__init__()
        self.aesthetic_target = aesthetic_target
        self.mode = mode
        extra = {"seed": weights_seed} if state_dict is None else {"state_dict": state_dict}
        self.model = models.CLIP("ViT-B-16", **extra)
        self.aesthetic_head = torch.nn.Linear(512, 10)
        if head_state_dict is not None:
            self.aesthetic_head.load_state_dict(head_state_dict)
        self.aesthetic_head.eval().requires_grad_(False)
        self.aesthetic_head.to(self.model.device)

    def forward(self, images):
        log_probs = self.aesthetic_head(self.model.encode_images(images))
        if self.mode == "logit":
            return -log_probs[..., self.aesthetic_target - 1].mean().mul(0.01)
        elif self.mode == "expected":
            expected_target = F.softmax(log_probs, dim=-1) * torch.arange(10).add(1).to(log_probs.device)
            return (expected_target - self.aesthetic_target).square().mean().mul(0.01)
        elif self.mode == "probability":
            return -F.softmax(log_probs, dim=-1)[..., self.aesthetic_target - 1].mean()
        else:
            raise ValueError(f"Unknown mode: {self.mode}")
