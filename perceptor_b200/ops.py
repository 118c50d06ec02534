"""Fine-grained torch-tensor wrappers over the per-kernel C-ABI entry points (tests, profiling, ncu runs).

The production path does not go through these: `pcg_guidance_fwd/_bwd` sequence the kernels natively.  Every
wrapper launches on torch's current stream, allocates outputs with torch, and raises on any non-zero return.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import native
from .resize_tables import ResizeTableCache

_p = native.ptr
bf16 = torch.bfloat16


def gemm(mode: int, a: torch.Tensor, b: torch.Tensor, bias=None, aux=None, act: int = native.ACT_QUICKGELU,
         bn: int = 0, k: int | None = None):
    """out[M,N] = epilogue(a[M,K] @ b[N,K]^T).  Returns out (and out2 for GEMM_BIAS_ACT)."""
    m, n = a.shape[0], b.shape[0]
    k = a.shape[1] if k is None else k
    out_dtype = torch.float32 if mode in (native.GEMM_RESID_F32, native.GEMM_F32) else bf16
    out = torch.empty((m, n), dtype=out_dtype, device=a.device)
    out2 = torch.empty((m, n), dtype=bf16, device=a.device) if mode == native.GEMM_BIAS_ACT else None
    lib = native.lib()
    args = (mode, act, m, n, k, _p(a), a.stride(0), _p(b), b.stride(0), _p(bias), _p(aux), _p(out), _p(out2), n,
            native.stream_ptr())
    rc = lib.pcg_gemm_bf16_bn(bn, *args) if bn else lib.pcg_gemm_bf16(*args)
    native.check(rc, "pcg_gemm_bf16")
    return (out, out2) if out2 is not None else out


def layernorm_fwd(x, gamma, beta):
    y = torch.empty(x.shape, dtype=bf16, device=x.device)
    native.check(native.lib().pcg_layernorm_fwd(_p(x), _p(gamma), _p(beta), _p(y), x.shape[0], x.shape[1],
                                                native.stream_ptr()), "pcg_layernorm_fwd")
    return y


def layernorm_bwd(dy, x, gamma, dx_io=None, dx_bf16=None):
    """dx += LN'(x)^T dy.  With `dx_io` (f32, updated in place) the bf16 copy is returned; with `dx_bf16` alone the
    bf16 tensor is the accumulator (updated in place and returned)."""
    dxb = dx_bf16 if dx_bf16 is not None else torch.empty(x.shape, dtype=bf16, device=x.device)
    native.check(native.lib().pcg_layernorm_bwd(_p(dy), _p(x), _p(gamma), _p(dx_io), _p(dxb), x.shape[0], x.shape[1],
                                                native.stream_ptr()), "pcg_layernorm_bwd")
    return dxb


def embed_fwd(patch_out, cls, pos, gamma, beta, n, tokens):
    d = patch_out.shape[1]
    v = torch.empty((n * tokens, d), dtype=torch.float32, device=patch_out.device)
    x0 = torch.empty_like(v)
    native.check(native.lib().pcg_embed_fwd(_p(patch_out), _p(cls), _p(pos), _p(gamma), _p(beta), _p(v), _p(x0), n,
                                            tokens, d, native.stream_ptr()), "pcg_embed_fwd")
    return v, x0


def embed_bwd(dx0, v, gamma, n, tokens):
    d = v.shape[1]
    out = torch.empty((n * (tokens - 1), d), dtype=bf16, device=v.device)
    f32 = dx0.dtype == torch.float32
    native.check(native.lib().pcg_embed_bwd(_p(dx0) if f32 else None, None if f32 else _p(dx0), _p(v), _p(gamma), _p(out),
                                            n, tokens, d, native.stream_ptr()), "pcg_embed_bwd")
    return out


def attn_fwd(qkv, n, tokens, heads):
    d = heads * 64
    out = torch.empty((n * tokens, d), dtype=bf16, device=qkv.device)
    lse = torch.empty((n, heads, tokens), dtype=torch.float32, device=qkv.device)
    native.check(native.lib().pcg_attn_fwd(_p(qkv), _p(out), _p(lse), n, tokens, heads, native.stream_ptr()),
                 "pcg_attn_fwd")
    return out, lse


def attn_bwd(qkv, out, d_out, lse, n, tokens, heads):
    d_qkv = torch.empty_like(qkv)
    delta = torch.empty(native.lib().pcg_attn_bwd_workspace_bytes(n, tokens, heads) // 4, dtype=torch.float32,
                        device=qkv.device)
    native.check(native.lib().pcg_attn_bwd(_p(qkv), _p(out), _p(d_out), _p(lse), _p(delta), _p(d_qkv), n, tokens, heads,
                                           native.stream_ptr()), "pcg_attn_bwd")
    return d_qkv


def attn_fwd_wide(qkv, n, tokens, heads):
    """Heads wider than 64 (80 / 88), stored padded to 128 columns: qkv [n*T, 3*heads*128]."""
    out = torch.empty((n * tokens, heads * 128), dtype=bf16, device=qkv.device)
    lse = torch.empty((n, heads, tokens), dtype=torch.float32, device=qkv.device)
    native.check(native.lib().pcg_attn_fwd_wide(_p(qkv), _p(out), _p(lse), n, tokens, heads, native.stream_ptr()),
                 "pcg_attn_fwd_wide")
    return out, lse


def attn_bwd_wide(qkv, out, d_out, lse, n, tokens, heads):
    d_qkv = torch.empty_like(qkv)
    delta = torch.empty(native.lib().pcg_attn_bwd_workspace_bytes(n, tokens, heads) // 4, dtype=torch.float32,
                        device=qkv.device)
    native.check(native.lib().pcg_attn_bwd_wide(_p(qkv), _p(out), _p(d_out), _p(lse), _p(delta), _p(d_qkv), n, tokens,
                                                heads, native.stream_ptr()), "pcg_attn_bwd_wide")
    return d_qkv


def attn_cls_fwd(qkv, n, tokens, heads, head_stride=64):
    """Class-token row only (last block): returns (out [n*T, heads*hs] with row n*T+0 written, lse [n, heads, T])."""
    out = torch.zeros((n * tokens, heads * head_stride), dtype=bf16, device=qkv.device)
    lse = torch.zeros((n, heads, tokens), dtype=torch.float32, device=qkv.device)
    native.check(native.lib().pcg_attn_cls_fwd(_p(qkv), _p(out), _p(lse), n, tokens, heads, head_stride,
                                               native.stream_ptr()), "pcg_attn_cls_fwd")
    return out, lse


def attn_cls_bwd(qkv, out, d_out, lse, n, tokens, heads, head_stride=64):
    d_qkv = torch.full_like(qkv, float("nan"))  # the kernel must write every element
    native.check(native.lib().pcg_attn_cls_bwd(_p(qkv), _p(out), _p(d_out), _p(lse), _p(d_qkv), n, tokens, heads,
                                               head_stride, native.stream_ptr()), "pcg_attn_cls_bwd")
    return d_qkv


def layernorm_fwd_rows(x, gamma, beta, y, rows, row_step):
    """LayerNorm of rows 0, row_step, 2*row_step, ... of x into the same rows of y (bf16, in place)."""
    native.check(native.lib().pcg_layernorm_fwd_rows(_p(x), _p(gamma), _p(beta), _p(y), rows, x.shape[1], row_step,
                                                     native.stream_ptr()), "pcg_layernorm_fwd_rows")
    return y


def layernorm_bwd_rows(dy, x, gamma, dx_bf16, rows, row_step):
    native.check(native.lib().pcg_layernorm_bwd_rows(_p(dy), _p(x), _p(gamma), None, _p(dx_bf16), rows, x.shape[1],
                                                     row_step, native.stream_ptr()), "pcg_layernorm_bwd_rows")
    return dx_bf16


def set_pooled_last_block(on: bool) -> bool:
    """Last transformer block on the class-token rows only (default) or in full; returns the previous setting."""
    return bool(native.lib().pcg_set_pooled_last_block(int(bool(on))))


def head_loss(x, ln_g, ln_b, proj, targets, tweights, n, tokens, scale=1.0, normalize=True, want_grad=True,
              d_enc=None):
    d, e = proj.shape
    dev = x.device
    loss = torch.zeros(1, dtype=torch.float32, device=dev)
    enc = torch.empty((n, e), dtype=torch.float32, device=dev)
    dx = torch.empty((n * tokens, d), dtype=torch.float32, device=dev) if want_grad else None
    dxb = torch.empty((n * tokens, d), dtype=bf16, device=dev) if want_grad else None
    m = 0 if targets is None else targets.shape[0]
    ws = torch.empty(native.lib().pcg_head_workspace_bytes(n, d, e) // 4, dtype=torch.float32, device=dev)
    native.check(native.lib().pcg_head_loss(_p(x), _p(ln_g), _p(ln_b), _p(proj), _p(targets), _p(tweights), n, tokens, d,
                                            e, m, float(scale), int(normalize), _p(loss), _p(enc), _p(d_enc), _p(dx),
                                            _p(dxb), _p(ws), native.stream_ptr()), "pcg_head_loss")
    return loss, enc, dx, dxb


class Sampler:
    """Per-kernel access to the cutout/resize/normalize sampler for one output size."""

    def __init__(self, out_size: int, patch: int, device, mean=(0.0, 0.0, 0.0), std=(1.0, 1.0, 1.0)):
        self.r, self.patch = out_size, patch
        self.kpatch = 3 * patch * patch
        self.kpad = (self.kpatch + 63) // 64 * 64
        self.device = torch.device(device)
        self.cache = ResizeTableCache(out_size)
        self.mean, self.std = native.host_floats(mean), native.host_floats(std)

    def _plan(self, rows5: np.ndarray, methods):
        dev = np.zeros((rows5.shape[0], native.CUT_STRIDE), dtype=np.int32)
        dev[:, :5] = rows5
        for i, (h, w) in enumerate(rows5[:, 3:5]):
            dev[i, 5] = self.cache.table_id(int(h), methods[i])
            dev[i, 6] = self.cache.table_id(int(w), methods[i])
        tabs = self.cache.device_tensors(self.device)
        tabs_c = native.ResizeTables(desc=tabs[0].data_ptr(), left=tabs[1].data_ptr(), weight=tabs[2].data_ptr(),
                                     inv=tabs[3].data_ptr(), n_desc=tabs[0].shape[0])
        return torch.from_numpy(dev).to(self.device), tabs_c, tabs, int(dev[:, 4].max())

    def forward(self, images, rows5, methods, want_patches=True, want_f32=True):
        table, tabs_c, keep, max_w = self._plan(np.asarray(rows5, dtype=np.int32), methods)
        n = table.shape[0]
        g = self.r // self.patch
        patches = torch.zeros((n * g * g, self.kpad), dtype=bf16, device=self.device) if want_patches else None
        out = torch.empty((n, 3, self.r, self.r), dtype=torch.float32, device=self.device) if want_f32 else None
        b, _, h, w = images.shape
        native.check(native.lib().pcg_sampler_fwd(_p(images), b, h, w, _p(table), n, C.byref(tabs_c), self.r, self.patch,
                                                  self.kpad, self.mean, self.std, _p(patches), _p(out), max_w,
                                                  native.stream_ptr()), "pcg_sampler_fwd")
        return patches, out

    def backward(self, images_shape, rows5, methods, d_patches=None, d_out=None):
        table, tabs_c, keep, max_w = self._plan(np.asarray(rows5, dtype=np.int32), methods)
        b, _, h, w = images_shape
        d_images = torch.zeros(images_shape, dtype=torch.float32, device=self.device)
        native.check(native.lib().pcg_sampler_bwd(_p(d_patches), _p(d_out), b, h, w, _p(table), table.shape[0],
                                                  C.byref(tabs_c), self.r, self.patch, self.kpad, self.std, _p(d_images),
                                                  max_w, native.stream_ptr()), "pcg_sampler_bwd")
        return d_images
