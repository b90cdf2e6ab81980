"""CPU tests of the drop-in boundary: the shared library builds for sm_100a, loads, exports every
symbol include/rs_retriever.h declares, rejects bad arguments without touching a device, and the
host/device layout function agrees with the oracle.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

from oracle import retrieval_oracle as o
from rag_serving_system_b200 import _cabi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rs_retriever.h")


def declared_symbols():
    txt = open(HEADER).read()
    return sorted(set(re.findall(r"RS_API\s+(?:const\s+char\*|int)\s+(rs_[a-z_]+)\s*\(", txt)))


def test_library_builds_and_loads():
    path = build.build()
    assert os.path.exists(path)
    lib = _cabi.load()
    assert lib.rs_version() == _cabi.ABI_VERSION == 2


def test_every_declared_symbol_is_exported():
    lib = _cabi.load()
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in rs_retriever.h but not exported"
    assert sorted(_cabi.EXPORTS) == syms


def test_struct_layouts_match_header():
    # sizes the C side checks via struct_size
    assert C.sizeof(_cabi.rs_config) == 48
    assert C.sizeof(_cabi.rs_stats) == 80


def test_build_targets_sm100a_only():
    assert "arch=compute_100a,code=sm_100a" in " ".join(build.NVCC_FLAGS)
    assert "-lineinfo" in build.NVCC_FLAGS


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_sass_shows_blackwell_paths():
    sass = subprocess.run(["cuobjdump", "-sass", build.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UBLKCP", "LDTM", "UTCBAR", "SYNCS"):
        assert mnemonic in sass, f"{mnemonic} missing: the fused kernel is not on the tcgen05/TMA path"
    assert "HMMA.16816" not in sass      # no legacy mma.sync path


def test_bad_arguments_fail_loudly_without_a_device():
    lib = _cabi.load()
    h = C.c_void_p()
    cfg = _cabi.rs_config(C.sizeof(_cabi.rs_config), 0, 1000, 0, 1000, 0, 0, 32, 16, 0)   # dim not multiple of 64
    assert lib.rs_create(C.byref(cfg), C.byref(h)) == -1
    assert b"multiple of 64" in lib.rs_last_error()
    cfg = _cabi.rs_config(4, 0, 1000, 0, 1024, 0, 0, 32, 16, 0)                           # wrong struct_size
    assert lib.rs_create(C.byref(cfg), C.byref(h)) == -1
    cfg = _cabi.rs_config(C.sizeof(_cabi.rs_config), 0, 0, 0, 1024, 0, 0, 32, 16, 0)      # no rows
    assert lib.rs_create(C.byref(cfg), C.byref(h)) == -1
    assert lib.rs_search(None, None, 1, 1, 0, 1, None, None) == -1
    assert lib.rs_destroy(None) == 0


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_cabi.RsError, match="no CPU fallback"):
        _cabi.Index(1000, 1024, "bf16")
    from rag_serving_system_b200 import CupyRetriever
    with pytest.raises(_cabi.RsError):
        CupyRetriever(np.zeros((10, 64), np.float32), [str(i) for i in range(10)])


def test_layout_header_matches_oracle(tmp_path):
    """rs_layout.h (shared by host and kernels) compiled with gcc == oracle.tile_layout_offsets."""
    src = tmp_path / "lay.c"
    src.write_text('#include "rs_layout.h"\n'
                   "unsigned long long lay(unsigned long long r, unsigned c, unsigned p, unsigned kc, unsigned pl)"
                   "{ return rs_corpus_offset(r, c, p, kc, pl); }\n")
    so = tmp_path / "lay.so"
    subprocess.check_call(["gcc", "-O1", "-shared", "-fPIC", "-I", os.path.join(build.CSRC), str(src), "-o", str(so)])
    lay = C.CDLL(str(so)).lay
    lay.restype = C.c_uint64
    lay.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    for planes, D, n in ((1, 128, 300), (2, 192, 140)):
        off = o.tile_layout_offsets(n, D, planes)
        rng = np.random.default_rng(0)
        for _ in range(2000):
            p, r, c = int(rng.integers(planes)), int(rng.integers(n)), int(rng.integers(D))
            assert lay(r, c, p, D // 64, planes) == off[p, r, c]
