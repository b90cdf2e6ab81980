"""TEST INFRASTRUCTURE: builds what the checker side needs beyond NumPy.

  build_native()  gcc-compiles oracle/synth_native.c (the C restatement of the synthetic corpus generator)
                  into oracle/_build/liboracle_native.so
  build_ref()     when /root/reference is present (the build container; it does not exist on the GPU box):
                  copies the UNMODIFIED reference module rag_service/core/retriever.py into oracle/_ref/ (git-ignored,
                  never committed, travels to the GPU box with the snapshot like a built .so) together with its
                  sha256, so that `bench.py --impl reference` can time the reference's own SimpleRetriever class
                  (retriever.py:57-166) instead of the oracle port.  Nothing under rag_serving_system_b200/ reads it.

Only tests/, __graft_entry__ and bench.py's CPU-baseline legs use these.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import shutil
import subprocess
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NATIVE_SRC = os.path.join(HERE, "synth_native.c")
NATIVE_LIB = os.path.join(HERE, "_build", "liboracle_native.so")
REF_SRC = "/root/reference/rag_service/core/retriever.py"
REF_DIR = os.path.join(HERE, "_ref")
REF_COPY = os.path.join(REF_DIR, "reference_retriever.py")


def build_native(force: bool = False) -> str | None:
    """Returns the library path, or None when gcc is unavailable (callers fall back to NumPy)."""
    if not force and os.path.exists(NATIVE_LIB) and os.path.getsize(NATIVE_LIB) > 0:
        return NATIVE_LIB
    gcc = shutil.which("gcc")
    if gcc is None:
        return None
    os.makedirs(os.path.dirname(NATIVE_LIB), exist_ok=True)
    tmp = NATIVE_LIB + f".tmp{os.getpid()}"
    res = subprocess.run([gcc, "-O2", "-fopenmp", "-shared", "-fPIC", NATIVE_SRC, "-o", tmp, "-lm"],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("gcc failed on oracle/synth_native.c:\n" + res.stderr)
    os.replace(tmp, NATIVE_LIB)
    return NATIVE_LIB


_native = None


def synth_rows_native(seed: int, row0: int, n_rows: int, D: int = 1024) -> np.ndarray | None:
    """Rows [row0, row0+n_rows) of the synthetic corpus as fp32, from the C restatement; None without gcc."""
    global _native
    if _native is None:
        path = build_native()
        if path is None:
            return None
        _native = C.CDLL(path)
        _native.oracle_synth_rows.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_int32, C.c_void_p]
        _native.oracle_synth_rows.restype = None
    out = np.empty((n_rows, D), dtype=np.float32)
    _native.oracle_synth_rows(seed, row0, n_rows, D, out.ctypes.data)
    return out


def build_ref() -> str | None:
    """Copy the unmodified reference module next to the oracle (build container only).  Returns the copy's path,
    the existing copy's path on a box without /root/reference, or None when neither exists."""
    if os.path.exists(REF_SRC):
        os.makedirs(REF_DIR, exist_ok=True)
        shutil.copyfile(REF_SRC, REF_COPY)
        with open(REF_COPY, "rb") as f:
            digest = hashlib.sha256(f.read()).hexdigest()
        with open(os.path.join(REF_DIR, "SHA256"), "w") as f:
            f.write(f"{digest}  rag_service/core/retriever.py (unmodified copy, made by oracle/build_oracle.py)\n")
    return REF_COPY if os.path.exists(REF_COPY) else None


def load_reference_module():
    """Import the unmodified reference module from oracle/_ref (None if it was never copied).  The module does
    `import cupy` at top level (retriever.py:9) and this image has no CuPy: an empty stand-in module satisfies
    the import; SimpleRetriever (the CPU class that is timed) never touches it."""
    path = build_ref()
    if path is None:
        return None
    import contextlib
    import importlib.util
    sys.modules.setdefault("cupy", types.ModuleType("cupy"))
    spec = importlib.util.spec_from_file_location("reference_retriever", path)
    mod = importlib.util.module_from_spec(spec)
    with contextlib.redirect_stdout(sys.stderr):      # the module prints its device choice at import (retriever.py:12-18)
        spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print("native:", build_native(force="--force" in sys.argv))
    print("ref:", build_ref())
