"""Drop-in retriever classes backed by the sm_100a C-ABI extension.

The call surface is the reference's (``rag_service/core/retriever.py``; citations below are
``retriever.py:<line>`` in the reference checkout):

    CupyRetriever(doc_embeddings, documents)          retriever.py:566
      .retrieve(query_emb, k=2)        -> List[str]   retriever.py:615
      .batch_retrieve(Q, ks)           -> List[List[str]]   retriever.py:686
      .documents / .N_A / .D                           retriever.py:589-591

so ``main.py:52`` and ``batch_processor.py:95`` work unchanged when this module stands in for
``rag_service.core.retriever``.  Added on top (north_star): ``search`` -> (indices, scores) and
``retrieve_with_scores`` -> (indices, scores, contexts).

There is NO CPU fallback: constructing a retriever without the built extension or without an
sm_100 GPU raises (the reference's ctor also raises when the device copy fails,
retriever.py:584-586).
"""
from __future__ import annotations

import json
import logging
import math
import threading
from operator import itemgetter
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _cabi

log = logging.getLogger(__name__)

_INGEST_ROWS = 1 << 16


def _auto_query_terms(q: np.ndarray, dtype: str) -> int:
    """1 if every query value is exactly representable in the corpus storage type (then a single
    tensor-core pass is exact), else 2 (hi + lo split, so fp32 queries keep their value)."""
    if dtype == "fp32":
        return 2
    if dtype == "bf16":
        return 1 if not np.any(q.view(np.uint32) & np.uint32(0xFFFF)) else 2
    with np.errstate(over="ignore"):
        return 1 if np.array_equal(q.astype(np.float16).astype(np.float32), q) else 2


def _fp32_scale_exp(max_abs: float) -> int:
    """Power-of-two pre-scale for fp32 storage (two fp16 planes): put max|x| in [2^3, 2^4) so the
    hi plane is far from fp16 overflow and the lo plane stays in the normal range."""
    if not (max_abs > 0.0) or not math.isfinite(max_abs):
        return 0
    return 3 - math.frexp(max_abs)[1] + 1


class B200Retriever:
    """Similarity + top-k over a device-resident corpus, one shard per device.

    metric="cosine" reproduces ``CupyRetriever.distance_cosine`` (retriever.py:594-611),
    metric="dot" reproduces ``SimpleRetriever`` (retriever.py:112).  ``dtype`` is the HBM storage
    type of the corpus ("fp32" keeps reference precision; "bf16"/"fp16" halve the bytes streamed).
    Ties are broken by the lower index (north_star contract; the reference is inconsistent).
    """

    def __init__(self, doc_embeddings: np.ndarray, documents: Sequence[str], *, dtype: str = "fp32",
                 metric: str = "cosine", devices: Optional[Sequence[int]] = None, max_batch: int = 256,
                 max_k: int = 128, query_terms: Optional[int] = None, source_dtype: Optional[str] = None):
        if not isinstance(doc_embeddings, np.ndarray) or doc_embeddings.ndim != 2:
            raise ValueError("doc_embeddings must be a 2-D numpy array")
        if dtype not in _cabi.DTYPES or metric not in _cabi.METRICS:
            raise ValueError(f"dtype in {list(_cabi.DTYPES)}, metric in {list(_cabi.METRICS)}")
        src_dtype = _source_dtype(doc_embeddings.dtype, source_dtype)
        if src_dtype is None:                                         # retriever.py:576-578
            doc_embeddings = doc_embeddings.astype(np.float32)
            src_dtype = "fp32"
        self._setup(doc_embeddings.shape[0], doc_embeddings.shape[1], documents, dtype, metric, devices,
                    max_batch, max_k, query_terms)
        scale_exp = 0
        if dtype == "fp32":
            scale_exp = _fp32_scale_exp(_max_abs_chunked(doc_embeddings, src_dtype))
        self._create_shards(scale_exp)
        try:
            for ix in self._shards:                                   # retriever.py:582 (H2D of the corpus)
                lo = ix.row_offset
                # chunk by chunk: works for np.memmap inputs far larger than host RAM
                for r in range(0, ix.n_rows, _INGEST_ROWS):
                    n = min(_INGEST_ROWS, ix.n_rows - r)
                    ix.load_rows(np.ascontiguousarray(doc_embeddings[lo + r: lo + r + n]), r, src_dtype)
                ix.finalize()
        except Exception:
            log.error("Failed to move document embeddings to the device", exc_info=True)   # retriever.py:584-586
            self.close()
            raise
        log.info("Initialized B200Retriever with %d documents (Dim=%d, dtype=%s, metric=%s, devices=%s)",
                 self.N_A, self.D, dtype, metric, self._devices)

    # ------------------------------------------------------------------ alternative constructors
    @classmethod
    def from_npy(cls, embeddings_path: str, contexts_path: Optional[str] = None, **kw) -> "B200Retriever":
        """main.py:20-23: a JSON list of strings + an .npy of embeddings.  The array is memory-mapped
        and ingested chunk by chunk (never materialised on the host).  fp32 / fp16 files load as
        they are; bf16 has no NumPy dtype - ml_dtypes writes it as '<V2' and a uint16 file needs
        ``source_dtype="bf16"``."""
        emb = np.load(embeddings_path, mmap_mode="r")
        docs = json.load(open(contexts_path)) if contexts_path else _LazyDocs(emb.shape[0])
        return cls(emb, docs, **kw)

    @classmethod
    def from_device_tensor(cls, emb, documents: Sequence[str], **kw) -> "B200Retriever":
        """Corpus already on a GPU (a torch CUDA tensor, fp32 / fp16 / bf16, contiguous): tiled in
        place by the ingest kernel, no host round trip.  Single device."""
        import torch
        assert emb.is_cuda and emb.dim() == 2 and emb.is_contiguous()
        src = {torch.float32: "fp32", torch.float16: "fp16", torch.bfloat16: "bf16"}[emb.dtype]
        self = cls.__new__(cls)
        dtype = kw.pop("dtype", "bf16" if src == "bf16" else ("fp16" if src == "fp16" else "fp32"))
        self._setup(emb.shape[0], emb.shape[1], documents, dtype, kw.pop("metric", "cosine"), [emb.device.index],
                    kw.pop("max_batch", 256), kw.pop("max_k", 128), kw.pop("query_terms", None))
        scale = _fp32_scale_exp(float(emb.abs().max())) if dtype == "fp32" else 0
        self._create_shards(scale)
        torch.cuda.synchronize(emb.device)
        self._shards[0].load_rows_ptr(emb.data_ptr(), emb.shape[0], src)
        self._shards[0].finalize()
        return self

    @classmethod
    def from_synthetic(cls, n_rows: int, dim: int = 1024, seed: int = 1234, documents=None, **kw) -> "B200Retriever":
        """Corpus generated ON DEVICE per shard (oracle: retrieval_oracle.synth_rows)."""
        self = cls.__new__(cls)
        docs = documents if documents is not None else _LazyDocs(n_rows)
        self._setup(n_rows, dim, docs, kw.pop("dtype", "bf16"), kw.pop("metric", "dot"), kw.pop("devices", None),
                    kw.pop("max_batch", 256), kw.pop("max_k", 128), kw.pop("query_terms", None))
        self._create_shards(4 if self._dtype == "fp32" else 0)   # unit-norm rows: |x| < 1
        for ix in self._shards:
            ix.fill_synthetic(seed)
            ix.finalize()
        return self

    # ------------------------------------------------------------------ internals
    def _setup(self, n, d, documents, dtype, metric, devices, max_batch, max_k, query_terms):
        self.documents = documents            # retriever.py:589 (kept on host, not copied)
        self.N_A = int(n)                     # retriever.py:590
        self.D = int(d)                       # retriever.py:591
        self._dtype, self._metric = dtype, metric
        self._devices = list(devices) if devices else [0]
        self._max_batch, self._max_k = int(max_batch), int(max_k)
        self._query_terms = query_terms
        self._shards: List[_cabi.Index] = []
        self._group: Optional[_cabi.Group] = None
        self._arena: Optional[_cabi.DocArena] = None   # byte arena of the documents, built on first use
        self._lock = threading.Lock()

    def _create_shards(self, scale_exp: int):
        G = len(self._devices)
        per = -(-self.N_A // G)
        for g, dev in enumerate(self._devices):
            lo = min(self.N_A, g * per)
            hi = min(self.N_A, lo + per)
            if hi > lo:
                self._shards.append(_cabi.Index(hi - lo, self.D, self._dtype, device=dev, row_offset=lo,
                                                max_batch=self._max_batch, max_k=self._max_k,
                                                scale_exp=scale_exp))
        if len(self._shards) > 1:
            # one process, several GPUs (main.py:52-55 constructs ONE retriever): the C ABI's group drives all
            # shards from the calling thread - one H2D + one fused kernel per GPU with the candidate exchange
            # inside the kernels (NVLink peer memory), one D2H; replayed from ONE multi-device CUDA graph
            self._group = _cabi.Group(self._shards)

    def close(self):
        if getattr(self, "_arena", None) is not None:
            self._arena.close()
            self._arena = None
        if getattr(self, "_group", None) is not None:
            self._group.close()
            self._group = None
        for ix in getattr(self, "_shards", []):
            ix.close()
        self._shards = []

    # ------------------------------------------------------------------ numeric API (added)
    def search(self, queries: np.ndarray, k: int, metric: Optional[str] = None,
               query_terms: Optional[int] = None) -> Tuple[np.ndarray, np.ndarray]:
        """(indices int64[B,k], scores float32[B,k]); rows past min(k, N) are (-1, -inf)."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if q.ndim != 2 or q.shape[1] != self.D:
            raise ValueError(f"queries must be [B, {self.D}]")
        if k <= 0:
            raise ValueError("k must be positive")
        if k > self._max_k:
            raise ValueError(f"k={k} exceeds max_k={self._max_k} given at construction")
        metric = metric or self._metric
        terms = query_terms or self._query_terms or _auto_query_terms(q, self._dtype)
        B = q.shape[0]
        backend = self._group if self._group is not None else self._shards[0]
        if B <= self._max_batch:                                  # the serving case: one call, fresh arrays
            with self._lock:
                return backend.search(q, k, metric, terms)
        out_i = np.full((B, k), -1, dtype=np.int64)
        out_s = np.full((B, k), -np.inf, dtype=np.float32)
        with self._lock:
            for b0 in range(0, B, self._max_batch):
                qb = q[b0:b0 + self._max_batch]
                i, s_ = backend.search(qb, k, metric, terms)
                out_i[b0:b0 + len(qb)] = i
                out_s[b0:b0 + len(qb)] = s_
        return out_i, out_s

    def search_tensor(self, queries, k: int, metric: Optional[str] = None, query_terms: Optional[int] = None):
        """Zero-copy hand-off from the embedder: ``queries`` is a float32 CUDA tensor
        [B, D] on the index's device; the search is enqueued on torch's current stream and the
        results come back as CUDA tensors (indices int64[B,k], scores float32[B,k]) without any
        host synchronisation.  batch_processor.py:91-92 could pass ``mean(dim=1)`` directly."""
        import torch
        if len(self._shards) != 1:
            raise ValueError("search_tensor drives a single-device index; use sharded.ShardedSearcher across GPUs")
        ix = self._shards[0]
        q = queries
        if q.dim() == 1:
            q = q[None, :]
        if not q.is_cuda or q.device.index != ix.device or q.shape[1] != self.D:
            raise ValueError(f"queries must be a CUDA tensor [B, {self.D}] on cuda:{ix.device}")
        q = q.contiguous().float()
        B = q.shape[0]
        if B > self._max_batch or k > self._max_k or k <= 0:
            raise ValueError("B / k outside the limits given at construction")
        terms = query_terms or self._query_terms or (2 if self._dtype == "fp32" else 1)
        oi = torch.empty((B, k), dtype=torch.int64, device=q.device)
        os_ = torch.empty((B, k), dtype=torch.float32, device=q.device)
        with self._lock:
            ix.search_device(q.data_ptr(), B, k, oi.data_ptr(), os_.data_ptr(), metric or self._metric, terms,
                             torch.cuda.current_stream(q.device).cuda_stream)
        return oi, os_

    def search_hidden_states(self, last_hidden_state, k: int, metric: Optional[str] = None):
        """batch_processor.py:89-97 without the host round trip: ``last_hidden_state`` is the embedder's output, a CUDA
        tensor [B, T, D] (fp32 / fp16 / bf16) on the index's device.  The mean over the sequence
        (``.mean(dim=1)``, batch_processor.py:91) is taken by ``rs_pool_mean_device`` straight into the query buffer and
        the search follows on torch's current stream; returns CUDA tensors (indices int64[B,k], scores float32[B,k],
        pooled queries float32[B,D])."""
        import torch
        if len(self._shards) != 1:
            raise ValueError("search_hidden_states drives a single-device index")
        ix = self._shards[0]
        h = last_hidden_state
        if not h.is_cuda or h.device.index != ix.device or h.dim() != 3 or h.shape[2] != self.D:
            raise ValueError(f"last_hidden_state must be a CUDA tensor [B, T, {self.D}] on cuda:{ix.device}")
        src = {torch.float32: "fp32", torch.float16: "fp16", torch.bfloat16: "bf16"}[h.dtype]
        h = h.contiguous()
        B, T = int(h.shape[0]), int(h.shape[1])
        if B > self._max_batch or k > self._max_k or k <= 0 or B <= 0 or T <= 0:
            raise ValueError("B / k outside the limits given at construction")
        stream = torch.cuda.current_stream(h.device).cuda_stream
        q = torch.empty((B, self.D), dtype=torch.float32, device=h.device)
        _cabi.pool_mean_device(ix.device, h.data_ptr(), src, B, T, self.D, q.data_ptr(), stream)
        terms = self._query_terms or 2       # pooled fp32 values are not representable in a 16-bit storage type
        oi = torch.empty((B, k), dtype=torch.int64, device=h.device)
        os_ = torch.empty((B, k), dtype=torch.float32, device=h.device)
        with self._lock:
            ix.search_device(q.data_ptr(), B, k, oi.data_ptr(), os_.data_ptr(), metric or self._metric, terms, stream)
        return oi, os_, q

    def retrieve_with_scores(self, queries: np.ndarray, k: int):
        """north_star triple: (indices, scores, contexts)."""
        idx, sc = self.search(queries, k)
        return idx, sc, [_gather(self.documents, row[row >= 0]) for row in idx]

    # ------------------------------------------------------------------ reference surface
    def retrieve(self, query_emb: np.ndarray, k: int = 2) -> List[str]:
        """retriever.py:615-683: one query, (D,) or (1,D); [] on any error, never raises."""
        if not isinstance(query_emb, np.ndarray):                         # :629-631
            log.error("retrieve expects query_emb as NumPy array, got %s", type(query_emb))
            return []
        if query_emb.ndim == 2 and query_emb.shape[0] == 1:               # :632-634
            query_emb = query_emb.flatten()
        elif query_emb.ndim != 1:                                         # :635-637
            log.error("retrieve expects a 1D query embedding or (1, dim), got shape %s", query_emb.shape)
            return []
        if query_emb.shape[0] != self.D:                                  # :638-640
            log.error("Query embedding dimension (%d) does not match document embedding dimension (%d)",
                      query_emb.shape[0], self.D)
            return []
        try:
            actual_k = min(k, self.N_A)                                   # :643
            if actual_k <= 0:                                             # :644-646
                log.warning("Requested k=%s results in actual_k=%s. Returning empty list.", k, actual_k)
                return []
            if actual_k > self._max_k:        # behavioural difference, INTEGRATION.md: the kernels keep <= max_k hits
                log.warning("k=%d exceeds max_k=%d; returning the best %d documents", actual_k, self._max_k, self._max_k)
                actual_k = self._max_k
            idx, _ = self.search(query_emb.reshape(1, self.D), actual_k)
            return _gather(self.documents, idx[0][idx[0] >= 0])           # :675
        except Exception as e:                                            # :681-683
            log.error("Error during retrieve: %s", e, exc_info=True)
            return []

    def batch_retrieve(self, query_embeddings: np.ndarray, ks: List[int]) -> List[List[str]]:
        """retriever.py:686-799: validation order and empty-list conventions preserved; the scores,
        per-row select and ordering (:742-769) run in the fused kernel at max_k, rows are then
        sliced to k_i (:776-790)."""
        if not isinstance(query_embeddings, np.ndarray):                  # :705-707
            log.error("batch_retrieve expects query_embeddings as NumPy array, got %s", type(query_embeddings))
            return [[] for _ in ks]
        if query_embeddings.ndim != 2:                                    # :708-710
            log.error("batch_retrieve expects a 2D query_embeddings array, got %d dimensions.", query_embeddings.ndim)
            return [[] for _ in ks]
        batch_size = query_embeddings.shape[0]
        if batch_size == 0:                                               # :711-713
            return []
        if not isinstance(ks, list) or batch_size != len(ks):             # :714-716
            log.error("Batch retrieve size mismatch: %d embeddings vs %s k values.", batch_size,
                      len(ks) if hasattr(ks, "__len__") else "?")
            return [[] for _ in range(batch_size)]
        if query_embeddings.shape[1] != self.D:                           # :717-719
            log.error("Batch query embedding dimension (%d) does not match document embedding dimension (%d)",
                      query_embeddings.shape[1], self.D)
            return [[] for _ in ks]
        if not all(isinstance(k_val, int) for k_val in ks):               # :720-722
            log.error("List ks must contain only integers.")
            return [[] for _ in ks]
        valid_ks = [k_val for k_val in ks if k_val > 0]                   # :725
        if not valid_ks:                                                  # :726-728
            log.warning("All k values in ks are non-positive. Returning list of empty lists.")
            return [[] for _ in ks]
        max_k = min(max(valid_ks), self.N_A)                              # :729-730
        if max_k <= 0:                                                    # :731-732
            return [[] for _ in ks]
        try:
            if max_k > self._max_k:
                # behavioural difference: the kernels keep at most max_k hits per query.  One
                # oversized request must not cost the other requests of its batch their contexts (the reference
                # serves any k <= N), so that row gets the best max_k documents and a warning.
                log.warning("batch_retrieve: k=%d exceeds max_k=%d; rows asking for more get the best %d documents",
                            max_k, self._max_k, self._max_k)
                max_k = self._max_k
            idx, _ = self.search(query_embeddings, max_k)                 # :736-773
            return _rows_to_docs(self.documents, idx, ks)                 # :776-790
        except Exception as e:                                            # :796-799
            log.error("Error during batch_retrieve: %s", e, exc_info=True)
            return [[] for _ in ks]

    # ------------------------------------------------------------------ contexts and prompts
    CONTEXT_SEPARATOR = "\n---\n"                                       # batch_processor.py:100

    def batch_retrieve_contexts(self, query_embeddings: np.ndarray, ks: List[int],
                                sep: Optional[str] = None) -> List[str]:
        """``[sep.join(docs) for docs in batch_retrieve(Q, ks)]`` (batch_processor.py:95-100) in one call: the hits are
        gathered and joined from a UTF-8 byte arena by the C ABI (rs_docs_join), so a batch costs B string objects
        instead of B*k lookups + B joins (25 600 lookups at B=256, k=100).  Same validation and error conventions
        as ``batch_retrieve``: a query whose row would be ``[]`` gets ``""``."""
        sep = self.CONTEXT_SEPARATOR if sep is None else sep
        if not isinstance(query_embeddings, np.ndarray) or query_embeddings.ndim != 2:
            return [sep.join(d) for d in self.batch_retrieve(query_embeddings, ks)]   # the reference's error paths
        B = query_embeddings.shape[0]
        if (B == 0 or not isinstance(ks, list) or B != len(ks) or query_embeddings.shape[1] != self.D
                or not all(isinstance(k_val, int) for k_val in ks)):
            return [sep.join(d) for d in self.batch_retrieve(query_embeddings, ks)]
        valid_ks = [k_val for k_val in ks if k_val > 0]
        max_k = min(max(valid_ks), self.N_A) if valid_ks else 0
        if max_k <= 0:
            return ["" for _ in ks]
        try:
            max_k = min(max_k, self._max_k)
            idx, _ = self.search(query_embeddings, max_k)
            if self._arena is None:
                if isinstance(self.documents, _LazyDocs):
                    return [sep.join(d) for d in _rows_to_docs(self.documents, idx, ks)]
                self._arena = _cabi.DocArena(self.documents)
            return self._arena.join(idx, ks, sep)
        except Exception as e:
            log.error("Error during batch_retrieve_contexts: %s", e, exc_info=True)
            return ["" for _ in ks]

    @staticmethod
    def assemble_prompts(queries: Sequence[str], contexts: Sequence[str]) -> List[str]:
        """The prompt template of batch_processor.py:103, applied to the joined contexts."""
        return [f"Context:\n{c}\n\nQuestion: {q}\n\nThe Answer to this question is: " for q, c in zip(queries, contexts)]

    def stats(self) -> List[dict]:
        return [ix.stats() for ix in self._shards]


def _rows_to_docs(documents, idx: np.ndarray, ks) -> List[List[str]]:
    """retriever.py:776-790 for a whole batch: row i -> documents of its first k_i valid (>= 0) indices, [] for
    k_i <= 0.  One ``tolist()`` for the index matrix and ``itemgetter`` for the lookups keep the per-query
    work out of NumPy (32 queries x 3 small-array operations cost more than the GPU search on a small
    corpus)."""
    is_list = isinstance(documents, list)
    B, K = idx.shape
    k0 = ks[0] if len(ks) else 0
    if is_list and 0 < k0 <= K and B * k0 > 1 and ks.count(k0) == B and int(idx[:, :k0].min()) >= 0:
        # the serving case (every request asks for the same k, corpus has at least k documents): ONE itemgetter call for
        # all B*k hits instead of one per query
        flat = itemgetter(*idx[:, :k0].ravel().tolist())(documents)
        return [list(flat[i * k0:(i + 1) * k0]) for i in range(B)]
    rows = idx.tolist()
    has_pad = bool((idx < 0).any())
    out: List[List[str]] = []
    for row, k_i in zip(rows, ks):
        if k_i <= 0:
            out.append([])
            continue
        if has_pad:
            row = [j for j in row if j >= 0]
        row = row[:k_i]
        n = len(row)
        if n == 0:
            out.append([])
        elif n == 1 or not is_list:
            out.append([documents[j] for j in row])
        else:
            out.append(list(itemgetter(*row)(documents)))
    return out


def _gather(documents, idx) -> List[str]:
    """documents[i] for i in idx (retriever.py:789).  ``itemgetter`` does the B*k lookups in C, which
    matters at k=100, B=256 (25 600 per call); the result is a fresh list."""
    n = len(idx)
    if n == 0:
        return []
    if n == 1 or not isinstance(documents, list):
        return [documents[int(i)] for i in idx]
    return list(itemgetter(*idx.tolist())(documents))


def _source_dtype(np_dtype, override: Optional[str]) -> Optional[str]:
    """Source element type the ingest kernel can read directly, or None (convert to fp32 first)."""
    if override is not None:
        if override not in _cabi.DTYPES:
            raise ValueError(f"source_dtype in {list(_cabi.DTYPES)}")
        want = 4 if override == "fp32" else 2
        if np_dtype.itemsize != want:
            raise ValueError(f"source_dtype={override} needs {want}-byte elements, array has {np_dtype}")
        return override
    if np_dtype == np.float32:
        return "fp32"
    if np_dtype == np.float16:
        return "fp16"
    if np_dtype.kind == "V" and np_dtype.itemsize == 2:
        return "bf16"
    return None


def _max_abs_chunked(emb: np.ndarray, src_dtype: str) -> float:
    m = 0.0
    for r in range(0, emb.shape[0], _INGEST_ROWS):
        c = np.asarray(emb[r:r + _INGEST_ROWS])
        if src_dtype == "bf16":
            c = (c.view(np.uint16).astype(np.uint32) << 16).view(np.float32)
        if c.size:
            m = max(m, float(np.abs(c).max()))
    return m


class _LazyDocs:
    """Stand-in document list for synthetic corpora: documents[i] == "doc-i" without storing N strings."""

    def __init__(self, n):
        self._n = n

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        i = int(i)
        if i < 0 or i >= self._n:
            raise IndexError(i)
        return f"doc-{i}"


class CupyRetriever(B200Retriever):
    """Name-compatible replacement for the class ``main.py:52`` constructs: cosine similarity,
    fp32-precision storage, positional (doc_embeddings, documents) constructor."""

    def __init__(self, doc_embeddings: np.ndarray, documents: List[str], **kw):
        kw.setdefault("metric", "cosine")
        kw.setdefault("dtype", "fp32")
        super().__init__(doc_embeddings, documents, **kw)


class SimpleRetriever(B200Retriever):
    """GPU-backed replacement for the reference's NumPy class (retriever.py:57-166): raw dot product."""

    def __init__(self, doc_embeddings: np.ndarray, documents: List[str], **kw):
        kw.setdefault("metric", "dot")
        kw.setdefault("dtype", "fp32")
        super().__init__(doc_embeddings, documents, **kw)
