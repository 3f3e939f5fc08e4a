"""Builds mmpl_b200/libmmpl_b200.so (the C-ABI library of include/mmpl_b200.h) with nvcc for sm_100a.

The library is built in-tree so that it travels with a snapshot of the repository; nvcc cross-compiles without a
GPU. There is no fallback: if the library is missing and cannot be built the import of the CUDA path fails loudly.

Staleness is decided by content, not by mtime: `source_id()` hashes every source, header and flag; the id is compiled
into the library (`mmpl_build_id()`) and written next to it (`libmmpl_b200.so.id`). A library whose id differs from the
sources in the tree is rebuilt, never loaded. Builds are serialised across processes with a file lock (torchrun starts
N ranks at once), objects are written under a unique name and renamed into place, and a failed build always raises.
"""
from __future__ import annotations

import contextlib
import fcntl
import hashlib
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
ID_PATH = PKG_DIR / "libmmpl_b200.so.id"
OBJ_DIR = REPO / "build" / "obj"

SOURCES = ["host_util.cu", "gemm_tcgen05.cu", "attention_tcgen05.cu", "attention_tcgen05_half.cu", "attention_dispatch.cu",
           "pointwise.cu", "sampler.cu", "vae_pointwise.cu", "api.cu"]
ID_CARRIER = "api.cu"  # compiled with -DMMPL_BUILD_ID

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-I", str(REPO / "include"), "-I", str(CSRC),
]


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: cannot build libmmpl_b200.so")
    return cand


def _headers():
    return sorted(list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [REPO / "include" / "mmpl_b200.h"])


def _digest(paths, extra: str = "") -> str:
    h = hashlib.sha256(extra.encode())
    for p in paths:
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()[:16]


def _flag_key() -> str:
    return " ".join(f for f in NVCC_FLAGS if not f.startswith("/"))  # paths differ between checkouts


def source_id() -> str:
    """Identity of what the library is built from: every .cu / header and the compiler flags."""
    return _digest(sorted(CSRC.glob('*.cu')) + _headers(), _flag_key())


def built_id() -> str:
    """Id recorded next to the library by the build that produced it ('' if there is none)."""
    if LIB_PATH.exists() and ID_PATH.exists():
        return ID_PATH.read_text().strip()
    return ""


def is_stale() -> bool:
    return built_id() != source_id()


@contextlib.contextmanager
def build_lock():
    """Exclusive inter-process lock for the object directory and the link step."""
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    with open(OBJ_DIR / ".lock", "w") as f:
        fcntl.flock(f, fcntl.LOCK_EX)
        try:
            yield
        finally:
            fcntl.flock(f, fcntl.LOCK_UN)


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a and link the shared library. Returns its path."""
    with build_lock():
        want = source_id()
        if not force and built_id() == want:  # another rank built it while this one waited for the lock
            return LIB_PATH
        nvcc = _nvcc()
        hdrs = _headers()

        def compile_one(src: str) -> Path:
            extra = [f'-DMMPL_BUILD_ID="{want}"'] if src == ID_CARRIER else []
            key = _digest(sorted(CSRC.glob('*.cu')) + hdrs if src.endswith('_half.cu') else [CSRC / src] + hdrs,
                          _flag_key() + " ".join(extra))
            obj = OBJ_DIR / f"{src[:-3]}.{key}.o"
            if obj.exists() and not force:
                return obj
            tmp = obj.with_suffix(f".{os.getpid()}.tmp")
            cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", str(CSRC / src), "-o", str(tmp)]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
            for old in OBJ_DIR.glob(f"{src[:-3]}.*.o"):
                old.unlink()
            os.replace(tmp, obj)
            return obj

        with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
            objs = list(ex.map(compile_one, SOURCES))
        tmp = LIB_PATH.with_suffix(f".so.{os.getpid()}.tmp")
        r = subprocess.run([nvcc, "-shared", "-o", str(tmp), *map(str, objs), "-cudart", "static"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if ID_PATH.exists():
            ID_PATH.unlink()  # never leave a new library next to an old id or the reverse
        os.replace(tmp, LIB_PATH)
        ID_PATH.write_text(want + "\n")
        return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
