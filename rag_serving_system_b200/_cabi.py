"""ctypes binding of include/rs_retriever.h (the C-ABI drop-in boundary).

There is no fallback: if the shared library is missing or a call fails, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

DTYPE_BF16, DTYPE_FP16, DTYPE_FP32 = 0, 1, 2
METRIC_DOT, METRIC_COSINE = 0, 1
DTYPES = {"bf16": DTYPE_BF16, "fp16": DTYPE_FP16, "fp32": DTYPE_FP32}
METRICS = {"dot": METRIC_DOT, "cosine": METRIC_COSINE}

EXPORTS = (
    "rs_version", "rs_last_error", "rs_create", "rs_destroy", "rs_load_rows", "rs_fill_synthetic",
    "rs_finalize", "rs_search", "rs_search_device", "rs_merge_device", "rs_get_rows",
    "rs_get_norms", "rs_get_image", "rs_get_stats", "rs_get_kernel_times",
    "rs_debug_plan", "rs_debug_timestamps", "rs_xchg_create", "rs_xchg_export", "rs_xchg_import", "rs_xchg_attach",
    "rs_xchg_destroy", "rs_search_exchange", "rs_search_exchange_host", "rs_check", "rs_set_option",
    "rs_get_rows_device", "rs_group_create", "rs_group_search", "rs_group_destroy",
    "rs_docs_create", "rs_docs_destroy", "rs_docs_join", "rs_pool_mean_device",
)
ABI_VERSION = 2          # include/rs_retriever.h: RS_ABI_VERSION
# switches of an index and their defaults (csrc/rs_api.cu: Options); rs_create reads the environment once
OPTION_DEFAULTS = {
    "RS_DOCM": 1, "RS_INLINE_PREP": 1, "RS_FUSED_TAIL": 1, "RS_DYN_SCHED": 1, "RS_SHARED_BOUND": 1,
    "RS_LISTS_GLOBAL": 1, "RS_GRAPH": 1, "RS_NO_EVENTS": 0, "RS_WIDE": 1, "RS_DMPAIR": 1, "RS_OVERLAP": 0,
    "RS_DEBUG_FLAGS": 0, "RS_DEBUG_TS": 0, "RS_QRESIDENT": -1, "RS_CPS": 0, "RS_SUB": 0, "RS_BUFC": 0,
    "RS_LEAN": -1, "RS_MMA_M": 0, "RS_NBUF": 0, "RS_QSTAGES": 0,
}


class RsError(RuntimeError):
    pass


class rs_config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("device", C.c_int32), ("n_rows", C.c_int64),
        ("row_offset", C.c_int64), ("dim", C.c_int32), ("dtype", C.c_int32),
        ("scale_exp", C.c_int32), ("max_batch", C.c_int32), ("max_k", C.c_int32),
        ("reserved", C.c_int32),
    ]


class rs_stats(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("sm_count", C.c_int32), ("grid", C.c_int32),
        ("stages", C.c_int32), ("smem_bytes", C.c_int32), ("passes", C.c_int32),
        ("launches", C.c_int32), ("query_terms", C.c_int32), ("q_resident", C.c_int32), ("sub", C.c_int32),
        ("nbuf", C.c_int32), ("mma_m", C.c_int32), ("overlapped", C.c_int32), ("doc_major", C.c_int32),
        ("bytes_streamed", C.c_int64),
        ("corpus_bytes", C.c_int64), ("total_launches", C.c_int64),
    ]


_lib = None


def lib_path() -> str:
    return _build.LIB_PATH


def load() -> C.CDLL:
    """Load librs_retriever.so (building it in-tree first if nvcc is present and it is stale)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path) or (_build.is_stale() and _build.have_nvcc()):
        path = _build.build()                   # no-op when the library is newer than every source
    if not os.path.exists(path):
        raise RsError(f"{path} is missing: run `python -m rag_serving_system_b200.build`; "
                      "there is no CPU fallback")
    lib = C.CDLL(path)
    vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
    lib.rs_version.restype = C.c_int
    if lib.rs_version() != ABI_VERSION:
        raise RsError(f"{path} has ABI version {lib.rs_version()}, this package expects {ABI_VERSION}: rebuild with "
                      "`python -m rag_serving_system_b200.build --force`")
    lib.rs_last_error.restype = C.c_char_p
    sigs = {
        "rs_create": [C.POINTER(rs_config), C.POINTER(vp)],
        "rs_destroy": [vp],
        "rs_load_rows": [vp, vp, i32, i64, i64],
        "rs_fill_synthetic": [vp, u64, i64],
        "rs_finalize": [vp],
        "rs_search": [vp, vp, i32, i32, i32, i32, vp, vp],
        "rs_search_device": [vp, vp, i32, i32, i32, i32, vp, vp, vp],
        "rs_merge_device": [i32, vp, vp, i32, i32, i32, i32, i64, i64, vp, vp, vp],
        "rs_get_kernel_times": [vp, vp, i32],
        "rs_debug_plan": [i32, i32, i32, i32, i32, vp],
        "rs_debug_timestamps": [vp, vp, i32],
        "rs_xchg_create": [i32, i32, i32, i32, i32, C.POINTER(vp)],
        "rs_xchg_export": [vp, vp],
        "rs_xchg_import": [vp, i32, vp],
        "rs_xchg_destroy": [vp],
        "rs_search_exchange": [vp, vp, vp, i32, i32, i32, i32, vp, vp, vp],
        "rs_search_exchange_host": [vp, vp, vp, i32, i32, i32, i32, vp, vp],
        "rs_xchg_attach": [vp, vp],
        "rs_check": [vp],
        "rs_set_option": [vp, C.c_char_p, i32],
        "rs_get_rows_device": [vp, i64, i64, vp, vp],
        "rs_group_create": [C.POINTER(vp), i32, C.POINTER(vp)],
        "rs_group_search": [vp, vp, i32, i32, i32, i32, vp, vp],
        "rs_group_destroy": [vp],
        "rs_docs_create": [vp, vp, i64, C.POINTER(vp)],
        "rs_docs_destroy": [vp],
        "rs_docs_join": [vp, vp, i32, i32, vp, C.c_char_p, i32, vp, i64, vp, vp],
        "rs_pool_mean_device": [i32, vp, i32, i32, i32, i32, vp, vp],
        "rs_get_rows": [vp, i64, i64, vp],
        "rs_get_norms": [vp, i64, i64, vp],
        "rs_get_image": [vp, i64, i64, vp],
        "rs_get_stats": [vp, C.POINTER(rs_stats)],
    }
    for name, args in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().rs_last_error().decode("utf-8", "replace")
        raise RsError(f"{what} failed (code {rc}): {msg}")


class Index:
    """Owner of one device shard (thin, explicit wrapper over the C ABI)."""

    def __init__(self, n_rows: int, dim: int, dtype: str = "bf16", device: int = 0, row_offset: int = 0,
                 max_batch: int = 256, max_k: int = 128, scale_exp: int = 0):
        self.lib = load()
        self.n_rows, self.dim, self.dtype, self.device = int(n_rows), int(dim), dtype, int(device)
        self.row_offset, self.max_batch, self.max_k = int(row_offset), int(max_batch), int(max_k)
        cfg = rs_config(C.sizeof(rs_config), device, n_rows, row_offset, dim, DTYPES[dtype], scale_exp,
                        max_batch, max_k, 0)
        h = C.c_void_p()
        check(self.lib.rs_create(C.byref(cfg), C.byref(h)), "rs_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.rs_destroy(self.h)
            self.h = None

    __del__ = close

    # ---- ingest
    def load_rows(self, rows: np.ndarray, row0: int = 0, src_dtype: str | None = None):
        rows = np.ascontiguousarray(rows)
        if src_dtype is None:
            if rows.dtype == np.float32:
                src_dtype = "fp32"
            elif rows.dtype == np.float16:
                src_dtype = "fp16"
            elif rows.dtype.itemsize == 2 and rows.dtype.kind == "V":
                src_dtype = "bf16"          # ml_dtypes' bfloat16 round-trips through .npy as '<V2'
            else:
                raise RsError(f"load_rows: pass float32/float16 rows or uint16 bits with src_dtype (got {rows.dtype})")
        assert rows.ndim == 2 and rows.shape[1] == self.dim
        check(self.lib.rs_load_rows(self.h, rows.ctypes.data, DTYPES[src_dtype], row0, rows.shape[0]), "rs_load_rows")

    def load_rows_ptr(self, ptr: int, nrows: int, src_dtype: str, row0: int = 0):
        """Rows behind a raw pointer: pageable / pinned host memory or DEVICE memory (tiled in place)."""
        check(self.lib.rs_load_rows(self.h, ptr, DTYPES[src_dtype], row0, nrows), "rs_load_rows")

    def fill_synthetic(self, seed: int, global_row0: int | None = None):
        g0 = self.row_offset if global_row0 is None else global_row0
        check(self.lib.rs_fill_synthetic(self.h, seed, g0), "rs_fill_synthetic")

    def finalize(self):
        check(self.lib.rs_finalize(self.h), "rs_finalize")

    # ---- search
    def search(self, q: np.ndarray, k: int, metric: str = "dot", query_terms: int = 1):
        q = np.ascontiguousarray(q, dtype=np.float32)
        B = q.shape[0]
        idx = np.empty((B, k), dtype=np.int64)
        sc = np.empty((B, k), dtype=np.float32)
        check(self.lib.rs_search(self.h, q.ctypes.data, B, k, METRICS[metric], query_terms, idx.ctypes.data,
                                 sc.ctypes.data), "rs_search")
        return idx, sc

    def search_device(self, q_ptr: int, B: int, k: int, out_idx_ptr: int, out_score_ptr: int, metric: str = "dot",
                      query_terms: int = 1, stream: int = 0):
        check(self.lib.rs_search_device(self.h, q_ptr, B, k, METRICS[metric], query_terms, out_idx_ptr,
                                        out_score_ptr, stream), "rs_search_device")

    def check(self):
        """Raise if a kernel of an earlier asynchronous search reported a time-out (rs_check)."""
        check(self.lib.rs_check(self.h), "rs_check")

    def set_option(self, name: str, value: int):
        """Change one RS_* switch after construction (the environment is read once, by rs_create)."""
        check(self.lib.rs_set_option(self.h, name.encode(), int(value)), "rs_set_option")

    def reload_env_options(self):
        """Re-read every RS_* switch from the environment (A/B tools flip them between runs on one index)."""
        for name, dflt in OPTION_DEFAULTS.items():
            v = os.environ.get(name, "")
            self.set_option(name, int(v) if v.strip() else dflt)

    # ---- introspection
    def get_rows_device(self, row0: int, nrows: int, out_ptr: int, stream: int = 0):
        """Rows [row0, row0+nrows) as float32 into DEVICE memory at out_ptr (asynchronous on `stream`)."""
        check(self.lib.rs_get_rows_device(self.h, row0, nrows, out_ptr, stream), "rs_get_rows_device")

    def get_rows(self, row0: int, nrows: int) -> np.ndarray:
        out = np.empty((nrows, self.dim), dtype=np.float32)
        check(self.lib.rs_get_rows(self.h, row0, nrows, out.ctypes.data), "rs_get_rows")
        return out

    def get_norms(self, row0: int, nrows: int) -> np.ndarray:
        out = np.empty(nrows, dtype=np.float32)
        check(self.lib.rs_get_norms(self.h, row0, nrows, out.ctypes.data), "rs_get_norms")
        return out

    def get_image(self, off: int, nbytes: int) -> np.ndarray:
        out = np.empty(nbytes, dtype=np.uint8)
        check(self.lib.rs_get_image(self.h, off, nbytes, out.ctypes.data), "rs_get_image")
        return out

    def kernel_times_ms(self, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.float32)
        check(self.lib.rs_get_kernel_times(self.h, out.ctypes.data, n), "rs_get_kernel_times")
        return out

    def debug_timestamps(self, n_ctas: int) -> np.ndarray:
        """uint64[n_ctas, 8] %globaltimer stamps of the last fused launch (needs RS_DEBUG_TS=1)."""
        out = np.zeros((n_ctas, 8), dtype=np.uint64)
        check(self.lib.rs_debug_timestamps(self.h, out.ctypes.data, n_ctas), "rs_debug_timestamps")
        return out

    def stats(self) -> dict:
        st = rs_stats()
        st.struct_size = C.sizeof(rs_stats)
        check(self.lib.rs_get_stats(self.h, C.byref(st)), "rs_get_stats")
        return {f: getattr(st, f) for f, _ in rs_stats._fields_}


class Exchange:
    """One rank's candidate exchange buffer (CUDA IPC, peer stores over NVLink)."""

    def __init__(self, device: int, world: int, rank: int, max_batch: int, max_k: int):
        self.lib = load()
        h = C.c_void_p()
        check(self.lib.rs_xchg_create(device, world, rank, max_batch, max_k, C.byref(h)), "rs_xchg_create")
        self.h, self.world, self.rank = h, world, rank

    def export_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        check(self.lib.rs_xchg_export(self.h, buf), "rs_xchg_export")
        return buf.raw

    def import_handle(self, peer_rank: int, handle: bytes):
        buf = C.create_string_buffer(bytes(handle), 64)
        check(self.lib.rs_xchg_import(self.h, peer_rank, buf), "rs_xchg_import")

    def search(self, index: "Index", q_ptr: int, B: int, k: int, out_idx_ptr: int, out_score_ptr: int,
               metric: str = "dot", query_terms: int = 1, stream: int = 0):
        check(self.lib.rs_search_exchange(index.h, self.h, q_ptr, B, k, METRICS[metric], query_terms, out_idx_ptr,
                                          out_score_ptr, stream), "rs_search_exchange")

    def search_host(self, index: "Index", q: np.ndarray, k: int, metric: str = "dot", query_terms: int = 1,
                    want_result: bool = True):
        """Host buffers in and out, blocking (rs_search_exchange_host).  want_result=False: this rank only takes
        part (no D2H) and gets (None, None)."""
        q = np.ascontiguousarray(q, dtype=np.float32)
        B = q.shape[0]
        if not want_result:
            check(self.lib.rs_search_exchange_host(index.h, self.h, q.ctypes.data, B, k, METRICS[metric], query_terms,
                                                   None, None), "rs_search_exchange_host")
            return None, None
        idx = np.empty((B, k), dtype=np.int64)
        sc = np.empty((B, k), dtype=np.float32)
        check(self.lib.rs_search_exchange_host(index.h, self.h, q.ctypes.data, B, k, METRICS[metric], query_terms,
                                               idx.ctypes.data, sc.ctypes.data), "rs_search_exchange_host")
        return idx, sc

    def close(self):
        if getattr(self, "h", None):
            self.lib.rs_xchg_destroy(self.h)
            self.h = None

    __del__ = close


class Group:
    """Single-process multi-GPU search over several ``Index`` shards on distinct devices (rs_group_*)."""

    def __init__(self, shards):
        self.lib = load()
        self.shards = list(shards)
        arr = (C.c_void_p * len(self.shards))(*[ix.h for ix in self.shards])
        h = C.c_void_p()
        check(self.lib.rs_group_create(arr, len(self.shards), C.byref(h)), "rs_group_create")
        self.h = h

    def search(self, q: np.ndarray, k: int, metric: str = "dot", query_terms: int = 1):
        q = np.ascontiguousarray(q, dtype=np.float32)
        B = q.shape[0]
        idx = np.empty((B, k), dtype=np.int64)
        sc = np.empty((B, k), dtype=np.float32)
        check(self.lib.rs_group_search(self.h, q.ctypes.data, B, k, METRICS[metric], query_terms, idx.ctypes.data,
                                       sc.ctypes.data), "rs_group_search")
        return idx, sc

    def close(self):
        if getattr(self, "h", None):
            self.lib.rs_group_destroy(self.h)
            self.h = None

    __del__ = close


def pool_mean_device(device: int, hidden_ptr: int, src_dtype: str, B: int, T: int, D: int, out_ptr: int, stream: int = 0):
    """Mean over the T positions of hidden [B,T,D] on `device` into fp32 [B,D], asynchronous on `stream` (rs_pool_mean_device)."""
    check(load().rs_pool_mean_device(device, hidden_ptr, DTYPES[src_dtype], B, T, D, out_ptr, stream), "rs_pool_mean_device")


class DocArena:
    """UTF-8 byte arena of the documents + native gather/join (rs_docs_*; host code of the C ABI, no GPU needed).

    ``join(idx, ks, sep)`` returns, per query, the documents of its first k_i valid hits joined with ``sep`` -
    what ``sep.join(batch_retrieve(...)[i])`` gives (retriever.py:776-790 + batch_processor.py:100) - with one
    memcpy per hit in C and ONE ``bytes.decode`` per query instead of B*k list lookups and B joins."""

    def __init__(self, documents):
        self.lib = load()
        enc = [d.encode("utf-8") for d in documents]
        off = np.zeros(len(enc) + 1, dtype=np.int64)
        if enc:
            np.cumsum(np.fromiter((len(e) for e in enc), dtype=np.int64, count=len(enc)), out=off[1:])
        arena = b"".join(enc)
        h = C.c_void_p()
        check(self.lib.rs_docs_create(arena, off.ctypes.data, len(enc), C.byref(h)), "rs_docs_create")
        self.h, self.n = h, len(enc)

    def join_bytes(self, idx: np.ndarray, ks, sep: str = "\n---\n"):
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        B, stride = idx.shape
        ks_a = np.ascontiguousarray(np.clip(np.asarray(ks, dtype=np.int64), 0, stride), dtype=np.int32)
        assert ks_a.shape == (B,)
        sep_b = sep.encode("utf-8")
        offs = np.empty(B + 1, dtype=np.int64)
        need = C.c_int64(0)
        check(self.lib.rs_docs_join(self.h, idx.ctypes.data, B, stride, ks_a.ctypes.data, sep_b, len(sep_b), None, 0,
                                    offs.ctypes.data, C.byref(need)), "rs_docs_join")
        buf = C.create_string_buffer(max(1, need.value))
        check(self.lib.rs_docs_join(self.h, idx.ctypes.data, B, stride, ks_a.ctypes.data, sep_b, len(sep_b), buf, need.value,
                                    offs.ctypes.data, C.byref(need)), "rs_docs_join")
        return buf.raw[:need.value], offs

    def join(self, idx: np.ndarray, ks, sep: str = "\n---\n"):
        raw, offs = self.join_bytes(idx, ks, sep)
        o = offs.tolist()
        return [raw[o[i]:o[i + 1]].decode("utf-8") for i in range(len(o) - 1)]

    def close(self):
        if getattr(self, "h", None):
            self.lib.rs_docs_destroy(self.h)
            self.h = None

    __del__ = close


def debug_plan(dim: int, dtype: str, B: int, k: int, query_terms: int = 1) -> dict:
    out = (C.c_int32 * 9)()
    check(load().rs_debug_plan(dim, DTYPES[dtype], B, k, query_terms, out), "rs_debug_plan")
    keys = ("passes", "per_pass", "bpad", "stages", "q_resident", "sub", "nbuf", "mma_m", "cps")
    return dict(zip(keys, list(out)))


def merge_device(device: int, scores_ptr: int, idx_ptr: int, n_lists: int, B: int, kc: int, k: int,
                 out_idx_ptr: int, out_score_ptr: int, stream: int = 0, score_list_stride: int = 0,
                 idx_list_stride: int = 0):
    check(load().rs_merge_device(device, scores_ptr, idx_ptr, n_lists, B, kc, k, score_list_stride,
                                 idx_list_stride, out_idx_ptr, out_score_ptr, stream), "rs_merge_device")
