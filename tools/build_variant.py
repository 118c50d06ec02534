"""Build an A/B variant of libpcg.so into tools/_build/libpcg_<name>.so with extra -D defines (or from another git
revision's csrc/), without touching the in-tree library.  Load it with PCG_LIBRARY=<path> (development aid only:
perceptor_b200/native.py honours the variable; the default stays perceptor_b200/libpcg.so).

    python tools/build_variant.py nobias -DPCG_EXP_NO_BIAS_PREFETCH
    python tools/build_variant.py old --rev HEAD~1
"""
import os
import shutil
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from perceptor_b200.build import NVCC_FLAGS, PRECISE_SOURCES  # noqa: E402


def main():
    name = sys.argv[1]
    rest = sys.argv[2:]
    rev = None
    if "--rev" in rest:
        i = rest.index("--rev")
        rev = rest[i + 1]
        rest = rest[:i] + rest[i + 2:]
    out_dir = ROOT / "tools" / "_build"
    out_dir.mkdir(exist_ok=True)
    work = Path(tempfile.mkdtemp(prefix="pcgvar_"))
    if rev:
        subprocess.run(f"git -C {ROOT} archive {rev} perceptor_b200/csrc include | tar -x -C {work}", shell=True, check=True)
        csrc, inc = work / "perceptor_b200" / "csrc", work / "include"
    else:
        csrc, inc = ROOT / "perceptor_b200" / "csrc", ROOT / "include"

    def one(src):
        obj = work / (src.stem + ".o")
        flags = [f for f in NVCC_FLAGS if not (src.name in PRECISE_SOURCES and f == "--use_fast_math") and f not in ("-Xptxas", "-v")]
        subprocess.run(["nvcc", *flags, *rest, "-I", str(inc), "-c", str(src), "-o", str(obj)], check=True)
        return str(obj)

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        objs = list(ex.map(one, sorted(csrc.glob("*.cu"))))
    out = out_dir / f"libpcg_{name}.so"
    subprocess.run(["nvcc", "-shared", "-o", str(out), *objs, "-gencode", "arch=compute_100a,code=sm_100a"], check=True)
    shutil.rmtree(work)
    print(out)


if __name__ == "__main__":
    main()
