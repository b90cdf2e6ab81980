"""In-tree build of the C-ABI extension (librs_retriever.so) for sm_100a.

``nvcc`` cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU
box with the repo snapshot.  Usage: ``python -m rag_serving_system_b200.build [--force]``.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "librs_retriever.so")
SOURCES = ["rs_api.cu", "rs_search.cu", "rs_ingest.cu"]
HEADERS = ["rs_kernels.h", "rs_layout.h", "rs_ptx.cuh", os.path.join("..", "..", "include", "rs_retriever.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",   # plain sm_100 rejects tcgen05.*
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-Xcompiler", "-pthread",
    "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build librs_retriever.so")


def have_nvcc() -> bool:
    try:
        _nvcc()
        return True
    except RuntimeError:
        return False


STAMP_PATH = LIB_PATH + ".stamp"


def _source_hash() -> str:
    """Content hash of everything the library is built from (mtimes do not survive a snapshot copy)."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for name in SOURCES + HEADERS:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH) or not os.path.exists(STAMP_PATH):
        return True
    with open(STAMP_PATH) as f:
        return f.read().strip() != _source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source of the extension into one shared library; returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS
    if verbose:
        cmd += ["-Xptxas", "-v"]
    tmp = LIB_PATH + f".tmp{os.getpid()}"       # several ranks may build at once: publish with an atomic rename
    cmd += ["-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode == 0:
        os.replace(tmp, LIB_PATH)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    with open(STAMP_PATH, "w") as f:
        f.write(_source_hash() + "\n")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
