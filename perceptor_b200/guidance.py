"""Host-side engine for the native guidance path: cutout plans, buffers, the two autograd Functions.

One `pcg_guidance_fwd` call runs sampler -> ViT -> head/loss on the current CUDA stream; one `pcg_guidance_bwd`
call runs the dgrad-only backward into the image gradient.  torch supplies device memory, streams and (for N>1)
the NCCL process group; nothing here computes on the CPU and there is no fallback.

Replaces CLIP.forward / OpenCLIP.encode_images of the reference (perceptor/losses/clip/clip.py:89-99,
perceptor/models/open_clip.py:109-123) and the autograd tape behind them.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from dataclasses import dataclass, replace

import numpy as np
import torch

from . import cutouts, native
from .resize_tables import CUBIC, LANCZOS3, ResizeTableCache, choose_method
from .vit import CLIP_MEAN, CLIP_STD, PackedWeights, VitShape


@dataclass
class CutPlan:
    """Cutouts of one call: public S0 rows, the device table the kernels read, and the shard of this rank."""

    rows: np.ndarray          # int32 [N_total, 4] (b, y0, x0, size) or [N_total, 5] (b, y0, x0, h, w)
    table: torch.Tensor       # int32 [n_local, 8] on device
    n_total: int
    n_local: int
    max_in_w: int
    tabs_c: native.ResizeTables
    tab_tensors: tuple        # keep-alive


class _GraphSlot:
    """Static buffers + the two captured CUDA graphs of one (shape, cutout count) configuration of the loss path.

    The guidance step is ~390 dependent launches; replaying them from a graph removes the launch gaps between them
    (ViT-L/14 x 128 cutouts: 53.7 -> 51.3 ms, ViT-B/32 x 256: 8.3 -> 7.8 ms on B200).  Everything a launch reads or
    writes lives in buffers owned by the slot; a call copies its inputs in (a few MB) and clones its outputs out."""

    def __init__(self, engine: "GuidanceEngine", images, plan: CutPlan, targets, tweights):
        dev = engine.device
        self.images = torch.empty_like(images)
        self.table = torch.empty_like(plan.table)
        self.targets = torch.empty_like(targets)
        self.tweights = torch.empty_like(tweights)
        self.loss_sum = torch.zeros(1, dtype=torch.float32, device=dev)
        self.d_images = torch.zeros_like(images)
        self.stash = torch.empty(engine.stash_bytes(plan.n_local), dtype=torch.uint8, device=dev)
        # the captured graphs bake this pointer in, so the slot owns its scratch memory: the engine-wide workspace of
        # the eager path may be reallocated by a later, larger call
        self.ws = torch.empty(engine.workspace_bytes(plan.n_local), dtype=torch.uint8, device=dev)
        self.fwd_graph: torch.cuda.CUDAGraph | None = None
        self.bwd_graph: torch.cuda.CUDAGraph | None = None
        self.calls_fwd = 0
        self.calls_bwd = 0
        # weak reference to the token of the forward whose activations the stash holds; it dies with that forward's
        # autograd node, so a loss that is never back-propagated does not pin the slot
        self.owner: weakref.ref | None = None
        self.launches_fwd = self.launches_bwd = 0

    def busy(self) -> bool:
        return self.owner is not None and self.owner() is not None


class _Token:
    __slots__ = ("__weakref__",)


class GuidanceEngine:
    """Owns packed weights, resize tables and scratch memory for one encoder on one device."""

    def __init__(self, shape: VitShape, state_dict, device, act: int, mean=CLIP_MEAN, std=CLIP_STD):
        self.shape = shape
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("the native guidance engine needs a CUDA device (there is no CPU fallback)")
        native.lib()  # fail loudly now if libpcg.so is missing
        self.weights = PackedWeights(shape, state_dict, self.device, act)
        self.tables = ResizeTableCache(shape.image_size)
        self._mean = native.host_floats(mean)
        self._std = native.host_floats(std)
        self._workspace: torch.Tensor | None = None
        self.launches_fwd = 0
        self.launches_bwd = 0
        # CUDA-graph replay of the loss path (PCG_CUDA_GRAPHS=0 turns it off); one slot is kept at a time
        self.use_graphs = os.environ.get("PCG_CUDA_GRAPHS", "1") != "0"
        self._slot_key = None
        self._slot: _GraphSlot | None = None
        self._prebuilt_side = 0

    # ------------------------------------------------------------------ plans
    def _method_for_size(self, s: int) -> int:
        return LANCZOS3 if s >= self.shape.image_size else CUBIC

    def _table_ids_for_sizes(self, sizes: np.ndarray) -> np.ndarray:
        """Resize-table id of every (square) cutout size: a lookup array filled on first use (the per-call dictionary this
        replaces cost 0.2 ms of host time per step, which an end-to-end step that reads its loss back cannot hide)."""
        lut = getattr(self, "_size_lut", None)
        top = int(sizes.max()) + 1
        if lut is None or lut.size < top:
            grown = np.full(max(top, 1024), -1, dtype=np.int32)
            if lut is not None:
                grown[:lut.size] = lut
            self._size_lut = lut = grown
        tid = lut[sizes]
        if (tid < 0).any():
            for s in np.unique(sizes[tid < 0]):
                lut[int(s)] = self.tables.table_id(int(s), self._method_for_size(int(s)))
            tid = lut[sizes]
        return tid

    def plan_cutouts(self, rows: np.ndarray, rank: int = 0, world: int = 1, b_offset: int | None = None) -> CutPlan:
        """rows: [N,4] square cutouts (b,y0,x0,size) or [N,5] boxes (b,y0,x0,h,w) of ALL ranks; this rank takes the
        contiguous shard `cutouts.shard_rows(N, rank, world)`.  `b_offset` (not None: image-sharded mode) re-bases the
        image index of the shard: the rank holds only its own images, global image b is local image b - b_offset, and
        `plan.rows` are the local, re-based rows (what the engine validates against the images it is given)."""
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        n_total = rows.shape[0]
        local = cutouts.local_rows(rows, rank, world, b_offset or 0)
        r = self.shape.image_size
        dev = np.zeros((local.shape[0], native.CUT_STRIDE), dtype=np.int32)
        if local.shape[0]:
            dev[:, :3] = local[:, :3]
            if rows.shape[1] == 4:
                sizes = local[:, 3]
                tid = self._table_ids_for_sizes(sizes)
                dev[:, 3] = dev[:, 4] = sizes
                dev[:, 5] = dev[:, 6] = tid
            else:
                for i, (h, w) in enumerate(local[:, 3:5]):
                    m = choose_method(int(h), int(w), r, r)
                    dev[i, 3], dev[i, 4] = h, w
                    dev[i, 5] = self.tables.table_id(int(h), m)
                    dev[i, 6] = self.tables.table_id(int(w), m)
        tab_tensors = self.tables.device_tensors(self.device)
        tabs_c = native.ResizeTables(desc=tab_tensors[0].data_ptr(), left=tab_tensors[1].data_ptr(),
                                     weight=tab_tensors[2].data_ptr(), inv=tab_tensors[3].data_ptr(),
                                     n_desc=tab_tensors[0].shape[0])
        table = torch.from_numpy(dev).pin_memory().to(self.device, non_blocking=True) if local.shape[0] else \
            torch.zeros((0, native.CUT_STRIDE), dtype=torch.int32, device=self.device)
        max_in_w = int(dev[:, 4].max()) if local.shape[0] else 1
        return CutPlan(rows if b_offset is None else local, table, n_total, local.shape[0], max_in_w, tabs_c, tab_tensors)

    def prebuild_tables(self, min_size: int, max_size: int) -> None:
        self.tables.ensure_sizes(range(min_size, max_size + 1), self._method_for_size)

    # ------------------------------------------------------------------ buffers
    def workspace_bytes(self, n: int) -> int:
        return native.lib().pcg_workspace_bytes(C.byref(self.weights.cfg), n)

    def _get_workspace(self, n: int) -> torch.Tensor:
        need = self.workspace_bytes(n)
        if self._workspace is None or self._workspace.numel() < need:
            self._workspace = None
            self._workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._workspace

    def stash_bytes(self, n: int) -> int:
        return native.lib().pcg_stash_bytes(C.byref(self.weights.cfg), n)

    def _args(self, images, plan: CutPlan, targets, tweights, loss_scale, workspace, stash, want_grad, loss_sum,
              enc_out, normalize, d_images=None, d_enc=None) -> native.GuidanceArgs:
        b, _, h, w = images.shape
        return native.GuidanceArgs(
            cfg=C.pointer(self.weights.cfg), w=C.pointer(self.weights.weights_c),
            images=native.ptr(images), B=b, H=h, W=w,
            cuts=native.ptr(plan.table), n_cut=plan.n_local, max_in_w=plan.max_in_w,
            tabs=C.pointer(plan.tabs_c), mean_host=self._mean, std_host=self._std,
            targets=native.ptr(targets), tweights=native.ptr(tweights),
            n_targets=0 if targets is None else targets.shape[0], loss_scale=float(loss_scale),
            workspace=native.ptr(workspace), workspace_bytes=workspace.numel(),
            stash=native.ptr(stash), stash_bytes=0 if stash is None else stash.numel(),
            want_grad=int(want_grad), loss_sum=native.ptr(loss_sum), enc_out=native.ptr(enc_out),
            normalize=int(normalize), d_images=native.ptr(d_images), d_enc=native.ptr(d_enc))

    # ------------------------------------------------------------------ forward / backward
    def forward(self, images: torch.Tensor, plan: CutPlan, targets, tweights, loss_scale: float, want_grad: bool,
                want_enc: bool, normalize: bool = True):
        """Returns (loss_sum [1] f32 or None, enc [n,E] f32 or None, stash or None)."""
        self._check_images(images, plan)
        lib = native.lib()
        n = plan.n_local
        loss_sum = torch.zeros(1, dtype=torch.float32, device=self.device) if targets is not None else None
        enc = torch.empty((n, self.shape.embed), dtype=torch.float32, device=self.device) if want_enc else None
        if n == 0:
            return loss_sum, enc, None
        ws = self._get_workspace(n)
        stash = torch.empty(self.stash_bytes(n), dtype=torch.uint8, device=self.device) if want_grad else None
        args = self._args(images, plan, targets, tweights, loss_scale, ws, stash, want_grad, loss_sum, enc, normalize)
        with torch.cuda.device(self.device):
            native.check(lib.pcg_guidance_fwd(C.byref(args), native.stream_ptr()), "pcg_guidance_fwd")
        self.launches_fwd = lib.pcg_last_launch_count()
        return loss_sum, enc, stash

    def backward(self, images_shape, plan: CutPlan, stash, targets, tweights, loss_scale: float,
                 normalize: bool = True, d_enc: torch.Tensor | None = None) -> torch.Tensor:
        """d(loss)/d(images), f32 [B,3,H,W] — the contribution of this rank's cutouts."""
        lib = native.lib()
        d_images = torch.zeros(images_shape, dtype=torch.float32, device=self.device)
        if plan.n_local == 0:
            return d_images
        ws = self._get_workspace(plan.n_local)
        args = self._args(d_images, plan, targets, tweights, loss_scale, ws, stash, True, None, None, normalize,
                          d_images=d_images, d_enc=d_enc)
        args.images = None
        with torch.cuda.device(self.device):
            native.check(lib.pcg_guidance_bwd(C.byref(args), native.stream_ptr()), "pcg_guidance_bwd")
        self.launches_bwd = lib.pcg_last_launch_count()
        return d_images

    # ------------------------------------------------------------------ CUDA-graph replay of the loss path
    def _slot_for(self, images, plan: CutPlan, targets, tweights, loss_scale: float):
        """The graph slot matching this call, (re)built when the configuration changes; None when graphs do not apply."""
        if not self.use_graphs or plan.n_local == 0 or torch.cuda.is_current_stream_capturing():
            return None
        # graphs bake the resize-table pointers in: build every table a crop of this image can need up front, so that
        # the device copy is uploaded once and never moves (table ids are append-only, older plans stay valid)
        side = int(max(images.shape[2], images.shape[3]))
        if side > self._prebuilt_side:
            self.prebuild_tables(1, side)
            self._prebuilt_side = side
        tabs = self.tables.device_tensors(self.device)
        key = (tuple(images.shape), plan.n_local, plan.n_total, tuple(targets.shape), float(loss_scale),
               tuple(t.data_ptr() for t in tabs))
        if self._slot is not None and self._slot.busy():
            # a forward is still waiting for its backward: its activations (and, if the configuration differs, its
            # whole slot: the pending backward replays graphs that point into it) must stay; this call runs eagerly
            return None
        if key != self._slot_key:
            self._slot = None  # frees the previous slot's stash before the new one is allocated
            self._slot_key = None
            self._slot = _GraphSlot(self, images, plan, targets, tweights)
            self._slot_key = key
        return self._slot

    def _slot_plan(self, slot: _GraphSlot, plan: CutPlan, width: int) -> CutPlan:
        # the sampler's launch geometry depends on the widest crop: fix it at the image width so one graph fits all;
        # the resize tables are the engine's complete, pinned-in-place set
        tabs = self.tables.device_tensors(self.device)
        tabs_c = native.ResizeTables(desc=tabs[0].data_ptr(), left=tabs[1].data_ptr(), weight=tabs[2].data_ptr(),
                                     inv=tabs[3].data_ptr(), n_desc=tabs[0].shape[0])
        return replace(plan, table=slot.table, max_in_w=int(width), tabs_c=tabs_c, tab_tensors=tabs)

    def forward_graphed(self, slot: _GraphSlot, images, plan: CutPlan, targets, tweights, loss_scale: float, token):
        """Loss forward (activations kept) through the slot.  The first call runs eagerly on the static buffers, the
        second captures, later ones replay."""
        self._check_images(images, plan)
        lib = native.lib()
        slot.images.copy_(images)
        slot.table.copy_(plan.table)
        slot.targets.copy_(targets)
        slot.tweights.copy_(tweights)
        splan = self._slot_plan(slot, plan, images.shape[3])
        ws = slot.ws

        def launch():
            slot.loss_sum.zero_()
            args = self._args(slot.images, splan, slot.targets, slot.tweights, loss_scale, ws, slot.stash, True,
                              slot.loss_sum, None, True)
            with torch.cuda.device(self.device):
                native.check(lib.pcg_guidance_fwd(C.byref(args), native.stream_ptr()), "pcg_guidance_fwd")
            slot.launches_fwd = lib.pcg_last_launch_count()

        if slot.fwd_graph is not None:
            slot.fwd_graph.replay()
        elif slot.calls_fwd >= 1:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                launch()
            slot.fwd_graph = graph
            graph.replay()
        else:
            launch()
        slot.calls_fwd += 1
        slot.owner = weakref.ref(token)
        self.launches_fwd = slot.launches_fwd
        return slot.loss_sum.clone()

    def backward_graphed(self, slot: _GraphSlot, plan: CutPlan, loss_scale: float) -> torch.Tensor:
        lib = native.lib()
        splan = self._slot_plan(slot, plan, slot.images.shape[3])
        ws = slot.ws

        def launch():
            slot.d_images.zero_()
            args = self._args(slot.d_images, splan, slot.targets, slot.tweights, loss_scale, ws, slot.stash, True, None,
                              None, True, d_images=slot.d_images)
            args.images = None
            with torch.cuda.device(self.device):
                native.check(lib.pcg_guidance_bwd(C.byref(args), native.stream_ptr()), "pcg_guidance_bwd")
            slot.launches_bwd = lib.pcg_last_launch_count()

        if slot.bwd_graph is not None:
            slot.bwd_graph.replay()
        elif slot.calls_bwd >= 1 and not torch.cuda.is_current_stream_capturing():
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                launch()
            slot.bwd_graph = graph
            graph.replay()
        else:
            launch()
        slot.calls_bwd += 1
        slot.owner = None
        self.launches_bwd = slot.launches_bwd
        return slot.d_images.clone()

    def _check_images(self, images: torch.Tensor, plan: CutPlan | None = None) -> None:
        if images.dim() != 4 or images.shape[1] != 3:
            raise ValueError(f"images must be [N,3,H,W], got {tuple(images.shape)}")
        if images.device != self.device or images.dtype != torch.float32 or not images.is_contiguous():
            raise ValueError("images must be contiguous float32 on the engine's CUDA device")
        if plan is not None:
            # the kernels index the images (and scatter-add into their gradient) straight from the cutout table
            cutouts.validate_rows(plan.rows, images.shape[0], images.shape[2], images.shape[3])


def _all_reduce_sum(t: torch.Tensor, group) -> None:
    if group is not None and torch.distributed.get_world_size(group) > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM, group=group)


class GuidanceLossFn(torch.autograd.Function):
    """loss = multiplier * mean_{n,m}(w_m * d(e_n, t_m)) over ALL ranks' cutouts; gradient into `images`."""

    @staticmethod
    def forward(ctx, images, engine: GuidanceEngine, plan: CutPlan, targets, tweights, multiplier, group,
                reduce_grad=True):
        """`reduce_grad`: the image gradient is summed over the group (cutout-sharded mode: every rank holds every
        image).  False in image-sharded mode, where a rank's cutouts only touch its own images: the step then has no
        data-path collective except the scalar loss."""
        want_grad = ctx.needs_input_grad[0]
        ctx.reduce_grad = bool(reduce_grad)
        scale = float(multiplier) / float(plan.n_total * targets.shape[0])
        slot = engine._slot_for(images, plan, targets, tweights, scale) if want_grad else None
        ctx.slot, ctx.token = slot, None
        if slot is not None:
            ctx.token = _Token()
            loss_sum, stash = engine.forward_graphed(slot, images, plan, targets, tweights, scale, ctx.token), None
        else:
            loss_sum, _, stash = engine.forward(images, plan, targets, tweights, scale, want_grad, False)
        _all_reduce_sum(loss_sum, group)
        ctx.engine, ctx.plan, ctx.stash, ctx.group, ctx.scale = engine, plan, stash, group, scale
        ctx.targets, ctx.tweights, ctx.images_shape = targets, tweights, tuple(images.shape)
        return loss_sum.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        if ctx.slot is not None:
            if ctx.slot.owner is None or ctx.slot.owner() is not ctx.token:
                raise RuntimeError("the activations of this forward were released (backward called twice?)")
            d_images = ctx.engine.backward_graphed(ctx.slot, ctx.plan, ctx.scale)
        else:
            d_images = ctx.engine.backward(ctx.images_shape, ctx.plan, ctx.stash, ctx.targets, ctx.tweights, ctx.scale)
        ctx.stash = None
        if ctx.reduce_grad:
            _all_reduce_sum(d_images, ctx.group)
        return d_images * grad_out, None, None, None, None, None, None, None


class EncodeImagesFn(torch.autograd.Function):
    """encodings [n_local, E] of this rank's cutouts, differentiable w.r.t. `images`."""

    @staticmethod
    def forward(ctx, images, engine: GuidanceEngine, plan: CutPlan, normalize):
        want_grad = ctx.needs_input_grad[0]
        _, enc, stash = engine.forward(images, plan, None, None, 0.0, want_grad, True, normalize)
        ctx.engine, ctx.plan, ctx.stash, ctx.normalize = engine, plan, stash, normalize
        ctx.images_shape = tuple(images.shape)
        return enc

    @staticmethod
    def backward(ctx, d_enc):
        d_enc = d_enc.contiguous().float()
        d_images = ctx.engine.backward(ctx.images_shape, ctx.plan, ctx.stash, None, None, 1.0, ctx.normalize, d_enc)
        ctx.stash = None
        return d_images, None, None, None
