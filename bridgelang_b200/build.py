"""build.py — compile csrc/*.cu for sm_100a and link the in-tree C-ABI library libbridgelang_b200.so.

nvcc cross-compiles without a GPU, so this runs in the CPU-only container too.  The CUDA runtime is linked
statically: the library has no dependency on the interpreter's torch build and exports only the `blb_*`
symbols declared in include/bridgelang_b200.h.
"""

from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
BUILD_DIR = PKG_DIR / "_build"
LIB_PATH = PKG_DIR / "libbridgelang_b200.so"
SOURCES = ["gemm_tcgen05.cu", "layernorm.cu", "attention.cu", "attention_tc.cu", "patch_embed.cu", "resize.cu", "decode_tail.cu", "capi.cu"]
HEADERS = [CSRC / "ptx.cuh", CSRC / "gemm.h", PKG_DIR.parent / "include" / "bridgelang_b200.h"]

# BLB_NVCC_EXTRA: extra nvcc flags for development builds (e.g. "-DBLB_ATTN_POLY_MASK=0"), part of the build digest
NVCC_FLAGS = os.environ.get("BLB_NVCC_EXTRA", "").split() + [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; libbridgelang_b200.so cannot be built")


def _digest() -> str:
    h = hashlib.sha256()
    for p in [CSRC / s for s in SOURCES] + HEADERS:
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile and link; no-op when the sources are unchanged since the last build."""
    BUILD_DIR.mkdir(exist_ok=True)
    stamp = BUILD_DIR / "stamp.txt"
    digest = _digest()
    if not force and LIB_PATH.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB_PATH
    nvcc = _nvcc()

    def compile_one(src: str) -> Path:
        obj = BUILD_DIR / (src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    link = [nvcc, "-shared", "-o", str(LIB_PATH), *map(str, objs), "-cudart", "static",
            "-gencode", "arch=compute_100a,code=sm_100a", "-Xlinker", "--no-undefined"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(digest)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
