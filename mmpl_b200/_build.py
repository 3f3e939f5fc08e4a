"""Builds mmpl_b200/libmmpl_b200.so (the C-ABI library of include/mmpl_b200.h) with nvcc for sm_100a.

The library is built in-tree so that it travels with a snapshot of the repository; nvcc
cross-compiles without a GPU. There is no fallback: if the library is missing and cannot be built the
import of the CUDA path fails loudly.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
REPO = PKG_DIR.parent
LIB_PATH = PKG_DIR / "libmmpl_b200.so"
OBJ_DIR = REPO / "build" / "obj"

SOURCES = ["host_util.cu", "gemm_tcgen05.cu", "attention_tcgen05.cu", "attention_tcgen05_half.cu", "attention_dispatch.cu", "pointwise.cu", "vae_pointwise.cu", "api.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-I", str(REPO / "include"), "-I", str(CSRC),
]



def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: cannot build libmmpl_b200.so")
    return cand


def _newest_source_mtime() -> float:
    files = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
    files.append(REPO / "include" / "mmpl_b200.h")
    return max(f.stat().st_mtime for f in files)


def is_stale() -> bool:
    return (not LIB_PATH.exists()) or LIB_PATH.stat().st_mtime < _newest_source_mtime()


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a and link the shared library. Returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = _nvcc()
    OBJ_DIR.mkdir(parents=True, exist_ok=True)

    def compile_one(src: str) -> Path:
        obj = OBJ_DIR / (src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = LIB_PATH.with_suffix(".so.tmp")
    cmd = [nvcc, "-shared", "-o", str(tmp), *map(str, objs), "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
