"""The drop-in boundary: libpcg.so loads, exports every symbol include/pcg.h declares, the binding covers them,
and the product never reaches into the oracle or a CPU fallback.  No compute calls here (no GPU needed)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "pcg.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pcg_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from perceptor_b200 import build, native

    path = build.build()
    assert path.exists()
    handle = ctypes.CDLL(str(path))
    names = declared_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(handle, name), f"{name} declared in include/pcg.h but not exported by libpcg.so"
        assert name in native.SIGNATURES, f"{name} has no ctypes signature in perceptor_b200/native.py"
    assert set(native.SIGNATURES) == set(names)
    lib = native.lib()
    assert lib.pcg_abi_version() == native.ABI_VERSION
    assert lib.pcg_last_error() is not None


def test_sizing_queries_need_no_gpu():
    from perceptor_b200 import native
    from perceptor_b200.vit import SHAPES

    s = SHAPES["ViT-L-14"]
    cfg = native.VitConfig(image_size=s.image_size, patch=s.patch, grid=s.grid, tokens=s.tokens, width=s.width,
                           layers=s.layers, heads=s.heads, mlp=s.mlp, embed=s.embed, kpatch=s.kpatch, kpad=s.kpad, act=0,
                           head_dim=s.head_dim)
    lib = native.lib()
    ws, st = lib.pcg_workspace_bytes(ctypes.byref(cfg), 128), lib.pcg_stash_bytes(ctypes.byref(cfg), 128)
    m = 128 * 257
    per_layer = m * 1024 * 4 + m * 3072 * 2 + m * 1024 * 2 + m * 4096 * 2
    assert st >= 24 * per_layer and st < 1.2 * (24 * per_layer + 26 * m * 1024 * 4)
    assert ws > 0 and lib.pcg_workspace_bytes(ctypes.byref(cfg), 0) == 0
    assert lib.pcg_stash_bytes(ctypes.byref(cfg), 256) > st
    # wide heads (ViT-H/14: head dim 80) are stored padded to 128 columns: the attention-side buffers grow accordingly
    assert lib.pcg_head_stride(64) == 64 and lib.pcg_head_stride(80) == 128 and lib.pcg_head_stride(88) == 128
    h = SHAPES["ViT-H-14"]
    cfg_h = native.VitConfig(image_size=h.image_size, patch=h.patch, grid=h.grid, tokens=h.tokens, width=h.width,
                             layers=h.layers, heads=h.heads, mlp=h.mlp, embed=h.embed, kpatch=h.kpatch, kpad=h.kpad, act=1,
                             head_dim=h.head_dim)
    mh = 4 * 257
    per_layer_h = mh * 1280 * 4 + mh * 3 * 2048 * 2 + mh * 2048 * 2 + mh * 5120 * 2
    assert lib.pcg_stash_bytes(ctypes.byref(cfg_h), 4) >= 32 * per_layer_h


def test_argument_errors_return_negative_codes_and_messages():
    from perceptor_b200 import native

    lib = native.lib()
    rc = lib.pcg_gemm_bf16(0, 0, 0, 64, 64, None, 64, None, 64, None, None, None, None, 64, None)
    assert rc < 0 and b"pcg_gemm_bf16" in lib.pcg_last_error()
    with pytest.raises(ValueError):
        native.check(rc, "gemm")
    rc = lib.pcg_attn_fwd(None, None, None, 1, 50, 12, None)
    assert rc < 0
    rc = lib.pcg_guidance_fwd(None, None)
    assert rc < 0


def test_product_never_imports_the_oracle_or_falls_back():
    pkg = ROOT / "perceptor_b200"
    for path in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        text = path.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{path} imports oracle/"
        assert "/root/reference" not in text, f"{path} reads the reference tree"
    gpu_side = [ROOT / "bench.py", ROOT / "__graft_entry__.py"] + list((ROOT / "tests").glob("test_gpu_*.py"))
    for path in gpu_side:
        if path.exists():
            assert "/root/reference" not in path.read_text(), f"{path} must not read /root/reference at run time"


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from perceptor_b200 import native

    monkeypatch.setattr(native, "_lib", None)
    monkeypatch.setattr(native, "LIB_PATH", tmp_path / "libpcg.so")
    with pytest.raises(native.NativeLibraryError):
        native.lib()
