"""ctypes binding of libpcg.so — the C-ABI CUDA library declared in include/pcg.h.

There is no CPU fallback: if the library is missing or fails to load, every entry point raises.  torch is used
only for device memory and streams; raw device pointers cross the ABI.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

import os

PKG_DIR = Path(__file__).resolve().parent
# PCG_LIBRARY: development aid for A/B runs of kernel variants (tools/build_variant.py); the product is the in-tree file
LIB_PATH = Path(os.environ["PCG_LIBRARY"]) if os.environ.get("PCG_LIBRARY") else PKG_DIR / "libpcg.so"

ACT_QUICKGELU, ACT_GELU = 0, 1
GEMM_BF16, GEMM_BIAS_ACT, GEMM_RESID_F32, GEMM_DACT, GEMM_F32 = 0, 1, 2, 3, 4
CUT_STRIDE = 8
ABI_VERSION = 4

_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)


class VitConfig(C.Structure):
    """pcg_vit_config"""

    _fields_ = [(n, C.c_int32) for n in (
        "image_size", "patch", "grid", "tokens", "width", "layers", "heads", "mlp", "embed", "kpatch", "kpad", "act",
        "head_dim")]


class LayerWeights(C.Structure):
    """pcg_layer_weights"""

    _fields_ = [(n, C.c_void_p) for n in (
        "ln1_g", "ln1_b", "ln2_g", "ln2_b",
        "w_qkv", "w_qkv_t", "w_out", "w_out_t", "w_fc", "w_fc_t", "w_proj", "w_proj_t",
        "b_qkv", "b_out", "b_fc", "b_proj")]


class VitWeights(C.Structure):
    """pcg_vit_weights"""

    _fields_ = [(n, C.c_void_p) for n in (
        "conv1", "conv1_t", "cls", "pos", "ln_pre_g", "ln_pre_b", "ln_post_g", "ln_post_b", "proj")] + [
        ("layers_host", C.POINTER(LayerWeights))]


class ResizeTables(C.Structure):
    """pcg_resize_tables"""

    _fields_ = [("desc", C.c_void_p), ("left", C.c_void_p), ("weight", C.c_void_p), ("inv", C.c_void_p),
                ("n_desc", C.c_int32)]


class GuidanceArgs(C.Structure):
    """pcg_guidance_args"""

    _fields_ = [
        ("cfg", C.POINTER(VitConfig)), ("w", C.POINTER(VitWeights)),
        ("images", C.c_void_p), ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("cuts", C.c_void_p), ("n_cut", C.c_int32), ("max_in_w", C.c_int32),
        ("tabs", C.POINTER(ResizeTables)),
        ("mean_host", _f32p), ("std_host", _f32p),
        ("targets", C.c_void_p), ("tweights", C.c_void_p), ("n_targets", C.c_int32),
        ("loss_scale", C.c_float),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("stash", C.c_void_p), ("stash_bytes", C.c_size_t),
        ("want_grad", C.c_int32),
        ("loss_sum", C.c_void_p), ("enc_out", C.c_void_p), ("normalize", C.c_int32),
        ("d_images", C.c_void_p), ("d_enc", C.c_void_p),
    ]


# name -> (restype, argtypes); must list every symbol include/pcg.h declares (tests check this against the header)
_vp, _i, _f = C.c_void_p, C.c_int, C.c_float
SIGNATURES = {
    "pcg_last_error": (C.c_char_p, []),
    "pcg_abi_version": (_i, []),
    "pcg_device_sm_count": (_i, []),
    "pcg_last_launch_count": (_i, []),
    "pcg_sampler_fwd": (_i, [_vp, _i, _i, _i, _vp, _i, C.POINTER(ResizeTables), _i, _i, _i, _f32p, _f32p, _vp, _vp, _i, _vp]),
    "pcg_sampler_bwd": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, C.POINTER(ResizeTables), _i, _i, _i, _f32p, _vp, _i, _vp]),
    "pcg_gemm_bf16": (_i, [_i, _i, _i, _i, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    "pcg_layernorm_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp]),
    "pcg_layernorm_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "pcg_layernorm_fwd_rows": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "pcg_layernorm_bwd_rows": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "pcg_attn_cls_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "pcg_attn_cls_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "pcg_set_pooled_last_block": (_i, [_i]),
    "pcg_embed_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "pcg_embed_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "pcg_head_stride": (_i, [_i]),
    "pcg_attn_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "pcg_attn_bwd_workspace_bytes": (C.c_size_t, [_i, _i, _i]),
    "pcg_attn_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "pcg_attn_fwd_wide": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "pcg_attn_bwd_wide": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "pcg_head_loss": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pcg_head_workspace_bytes": (C.c_size_t, [_i, _i, _i]),
    "pcg_workspace_bytes": (C.c_size_t, [C.POINTER(VitConfig), _i]),
    "pcg_stash_bytes": (C.c_size_t, [C.POINTER(VitConfig), _i]),
    "pcg_guidance_fwd": (_i, [C.POINTER(GuidanceArgs), _vp]),
    "pcg_guidance_bwd": (_i, [C.POINTER(GuidanceArgs), _vp]),
    "pcg_affine2": (_i, [_vp, _vp, _vp, _vp, _i, C.c_longlong, _f, _vp]),
    "pcg_clamp_with_grad": (_i, [_vp, _vp, _vp, C.c_longlong, _f, _f, _vp]),
    "pcg_profile_enable": (_i, [_i]),
    "pcg_profile_collect": (_i, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]),
}
PROF_KINDS = ("gemm", "attn_fwd", "attn_bwd", "layernorm", "sampler_fwd", "sampler_bwd", "embed", "head")


def profile_collect() -> dict:
    """{family: {"ms": .., "work": .., "count": ..}} since the last collect (pcg_profile_enable(1) first)."""
    n = len(PROF_KINDS)
    ms, work, cnt = (C.c_double * n)(), (C.c_double * n)(), (C.c_int * n)()
    check(lib().pcg_profile_collect(ms, work, cnt), "pcg_profile_collect")
    return {k: {"ms": ms[i], "work": work[i], "count": cnt[i]} for i, k in enumerate(PROF_KINDS)}
# test hook, not part of the public header
_EXTRA = {"pcg_gemm_bf16_bn": (_i, [_i, _i, _i, _i, _i, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp]),
          "pcg_gemm_set_variant": (_i, [_i]), "pcg_attn_set_legacy": (_i, [_i]), "pcg_attn_set_persist": (_i, [_i]), "pcg_attn_set_split": (_i, [_i]), "pcg_attn_set_trace": (_i, [_vp]), "pcg_sampler_set_scalar": (_i, [_i])}

_lib = None


class NativeLibraryError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load libpcg.so once.  Raises NativeLibraryError (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise NativeLibraryError(
            f"{LIB_PATH} is missing: build it with `python -m perceptor_b200.build` "
            "(there is no CPU or PyTorch fallback for the guidance path)")
    try:
        handle = C.CDLL(str(LIB_PATH))
    except OSError as e:  # pragma: no cover - depends on the environment
        raise NativeLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in {**SIGNATURES, **_EXTRA}.items():
        if name in _EXTRA and os.environ.get("PCG_LIBRARY") and not hasattr(handle, name):
            continue  # an older A/B variant may lack a newer test hook
        fn = getattr(handle, name)
        fn.restype = res
        fn.argtypes = args
    if handle.pcg_abi_version() != ABI_VERSION:
        raise NativeLibraryError(f"libpcg.so ABI {handle.pcg_abi_version()} != binding ABI {ABI_VERSION}; rebuild")
    _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    """Map the C-ABI return code onto the reference's error behaviour (ValueError / RuntimeError)."""
    if rc == 0:
        return
    msg = lib().pcg_last_error().decode("utf-8", "replace")
    if rc < 0:
        raise ValueError(f"{what}: {msg}")
    raise RuntimeError(f"{what}: CUDA error {rc}: {msg}")


def ptr(t: torch.Tensor | None) -> int | None:
    """Device pointer of a contiguous CUDA tensor (None passes NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise ValueError("the native guidance path needs CUDA tensors (there is no CPU fallback)")
    if not t.is_contiguous():
        raise ValueError("the native guidance path needs contiguous tensors")
    return t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def host_floats(values) -> C.Array:
    return (C.c_float * len(values))(*[float(v) for v in values])
