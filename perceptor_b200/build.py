"""Build libpcg.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

`python -m perceptor_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles without a GPU.  The built
library is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
BUILD_DIR = PKG_DIR / "_build"
LIB_PATH = PKG_DIR / "libpcg.so"

NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-O3",
    "-std=c++17",
    "--use_fast_math",
    "-Xcompiler",
    "-fPIC,-O3,-Wall,-Wno-unused-function",
    "-Xptxas",
    "-v",
]
# --use_fast_math would turn sqrtf/divisions in LayerNorm and the loss head into approximations; those kernels
# are HBM/latency bound, so they are compiled without it.
PRECISE_SOURCES = {"rowwise.cu", "sampler.cu"}


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; libpcg.so cannot be built")
    return cand


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG_DIR.parent / "include" / "pcg.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_one(src: Path) -> tuple[Path, str]:
    obj = BUILD_DIR / (src.stem + ".o")
    flags = [f for f in NVCC_FLAGS if not (src.name in PRECISE_SOURCES and f == "--use_fast_math")]
    cmd = [_nvcc(), *flags, "-I", str(PKG_DIR.parent / "include"), "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    return obj, r.stderr


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ for sm_100a and link libpcg.so.  Skips when sources are unchanged."""
    stamp = BUILD_DIR / "digest.txt"
    digest = _digest()
    if not force and LIB_PATH.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB_PATH
    BUILD_DIR.mkdir(exist_ok=True)
    with concurrent.futures.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        results = list(ex.map(_compile_one, _sources()))
    (BUILD_DIR / "ptxas.log").write_text("\n".join(log for _, log in results))
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    objs = [str(o) for o, _ in results]
    cmd = [_nvcc(), "-shared", "-o", str(LIB_PATH), *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(digest)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
