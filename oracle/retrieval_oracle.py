"""CPU oracle for the similarity + top-k hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``rag_serving_system_b200``) never imports it and has no CPU fallback.

Every function restates, in NumPy, the arithmetic of the reference file
``rag_service/core/retriever.py`` (citations are ``retriever.py:<line>`` relative to
``/root/reference``).  The restatement is pinned against the *imported, unmodified*
reference ``SimpleRetriever`` (raw dot) AND the imported, unmodified live class ``CupyRetriever``
(cosine) by ``oracle/make_golden.py`` (run in the build container, where ``/root/reference``
exists); the vectors it wrote live in ``tests/golden/`` and are re-checked by
``tests/test_oracle.py`` on every run.  The reference ships no tests of its own.
``CupyRetriever`` holds no arithmetic of its own - every operation is a call into CuPy, which is
not vendored, not pinned (Dockerfile:12; ``#cupy==13.4.1`` at requirements.txt:84) and not
installable here - so the generator runs the reference class with ``oracle/cupy_numpy_shim.py``
registered as ``cupy``: the reference's own control flow and call sequence, NumPy's fp32
implementation of each call.  That pins the cosine formula, its operation order, ``max_k``
handling, slicing and string gather; what it cannot pin is cuBLAS' summation order inside SGEMM
(covered by the 1e-5 relative score bar) and CuPy's tie order (goldens are tie-free).
"""
from __future__ import annotations

import logging
from typing import List, Sequence, Tuple

import numpy as np

# ----------------------------------------------------------------------------------------------
# dtype helpers (the kernel stores the corpus as bf16 / fp16 / split-fp16; the reference always
# computes in fp32 on whatever values it is handed, retriever.py:71-73,576-578)
# ----------------------------------------------------------------------------------------------

DTYPE_BF16 = 0
DTYPE_FP16 = 1
DTYPE_FP32 = 2
DTYPE_NAMES = {"bf16": DTYPE_BF16, "fp16": DTYPE_FP16, "fp32": DTYPE_FP32}


def round_to_bf16(x: np.ndarray) -> np.ndarray:
    """fp32 -> nearest-even bf16 -> fp32 (values exactly representable in bf16)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).reshape(np.shape(x))


def bf16_bits(x: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 bit patterns (uint16), round-to-nearest-even."""
    return (round_to_bf16(x).view(np.uint32) >> 16).astype(np.uint16)


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint32) << 16).view(np.float32)


def round_to_fp16(x: np.ndarray) -> np.ndarray:
    return np.asarray(x, dtype=np.float32).astype(np.float16).astype(np.float32)


def round_to_dtype(x: np.ndarray, dtype: int) -> np.ndarray:
    """The values the kernel sees for a corpus / query stored in ``dtype`` (as fp32)."""
    if dtype == DTYPE_BF16:
        return round_to_bf16(x)
    if dtype == DTYPE_FP16:
        return round_to_fp16(x)
    return np.asarray(x, dtype=np.float32)


FP16_SPLIT_SHIFT = 11  # lo plane is stored scaled by 2**11


def split_fp16(x: np.ndarray, scale_exp: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """fp32 -> (hi, lo) fp16 planes with  x * 2**scale_exp ~= hi + lo * 2**-11.

    This is the on-device storage of an fp32 corpus (4 bytes/element, same as fp32, but
    directly consumable by kind::f16 tensor-core MMAs).  All steps are exact fp32 operations
    except the two fp16 roundings.
    """
    xs = np.ldexp(np.asarray(x, dtype=np.float32), scale_exp).astype(np.float32)
    hi = xs.astype(np.float16)
    r = (xs - hi.astype(np.float32)).astype(np.float32)  # exact in fp32
    lo = np.ldexp(r, FP16_SPLIT_SHIFT).astype(np.float32).astype(np.float16)
    return hi, lo


def unsplit_fp16(hi: np.ndarray, lo: np.ndarray, scale_exp: int = 0) -> np.ndarray:
    """Value represented by a split pair, as float64 (exact)."""
    v = hi.astype(np.float64) + np.ldexp(lo.astype(np.float64), -FP16_SPLIT_SHIFT)
    return np.ldexp(v, -scale_exp)


def split_bf16(x: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """fp32 -> (hi, lo) both bf16-representable fp32 arrays, x ~= hi + lo (16 mantissa bits)."""
    x = np.asarray(x, dtype=np.float32)
    hi = round_to_bf16(x)
    lo = round_to_bf16((x - hi).astype(np.float32))
    return hi, lo


# ----------------------------------------------------------------------------------------------
# similarity (retriever.py:112 raw dot; retriever.py:594-611 cosine)
# ----------------------------------------------------------------------------------------------

def scores_dot(Q: np.ndarray, D: np.ndarray, accum=np.float32) -> np.ndarray:
    """``sims = doc_embeddings @ query_emb`` for every query (retriever.py:112), as [B, N]."""
    Q = np.asarray(Q, dtype=accum)
    D = np.asarray(D, dtype=accum)
    return (Q @ D.T).astype(accum)


def scores_cosine(Q: np.ndarray, D: np.ndarray, epsilon: float = 1e-8, accum=np.float32) -> np.ndarray:
    """clip( (X.Y^T) / (|X| |Y|^T + eps), -1, 1 )   (retriever.py:600-609), as [B, N]."""
    X = np.asarray(Q, dtype=accum)
    Y = np.asarray(D, dtype=accum)
    dot = X @ Y.T                                                  # :600
    nx = np.linalg.norm(X, axis=1, keepdims=True).astype(accum)    # :602
    ny = np.linalg.norm(Y, axis=1, keepdims=True).astype(accum)    # :603
    prod = (nx @ ny.T).astype(accum)                               # :605
    cos = dot / (prod + accum(epsilon))                            # :607
    return np.clip(cos, -1.0, 1.0).astype(accum)                   # :609


def scores(Q, D, metric: str = "dot", accum=np.float32) -> np.ndarray:
    if metric == "dot":
        return scores_dot(Q, D, accum)
    if metric == "cosine":
        return scores_cosine(Q, D, accum=accum)
    raise ValueError(metric)


# ----------------------------------------------------------------------------------------------
# top-k (retriever.py:118-120, :757-763) with the north_star tie contract: ties -> lower index
# ----------------------------------------------------------------------------------------------

def topk_rows(S: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Per row: the k best (score desc, index asc).  Pads with (-1, -inf) when N < k
    (retriever.py:767-769 pads indices with -1)."""
    S = np.asarray(S)
    B, N = S.shape
    kk = min(k, N)
    idx = np.full((B, k), -1, dtype=np.int64)
    val = np.full((B, k), -np.inf, dtype=np.float32)
    if kk > 0:
        # stable argsort of -S gives descending score with ascending index inside equal groups
        order = np.argsort(-S, axis=1, kind="stable")[:, :kk]
        idx[:, :kk] = order
        val[:, :kk] = np.take_along_axis(S, order, axis=1)
    return idx, val


def search(Q, D, k: int, metric: str = "dot", accum=np.float32):
    """The contract of the CUDA path: (indices int64[B,k], scores float32[B,k])."""
    return topk_rows(scores(Q, D, metric, accum), k)


def merge_candidates(cand_scores: np.ndarray, cand_idx: np.ndarray, k: int):
    """Merge G candidate lists per query ([G, B, kc] each) into the global top-k, ordering by
    (score desc, global index asc); entries with index < 0 are padding.  This is the
    restatement of the multi-CTA and multi-GPU merge steps (north_star (b), (c))."""
    cs = np.asarray(cand_scores, dtype=np.float32)
    ci = np.asarray(cand_idx, dtype=np.int64)
    G, B, kc = cs.shape
    cs = np.transpose(cs, (1, 0, 2)).reshape(B, G * kc)
    ci = np.transpose(ci, (1, 0, 2)).reshape(B, G * kc)
    out_i = np.full((B, k), -1, dtype=np.int64)
    out_s = np.full((B, k), -np.inf, dtype=np.float32)
    for b in range(B):
        valid = ci[b] >= 0
        s, i = cs[b][valid], ci[b][valid]
        order = np.lexsort((i, -s.astype(np.float64)))[:k]
        out_i[b, : len(order)] = i[order]
        out_s[b, : len(order)] = s[order]
    return out_i, out_s


def shard_range(N: int, G: int, g: int) -> Tuple[int, int]:
    """Contiguous row shards: GPU g owns [g*ceil(N/G), min(N,(g+1)*ceil(N/G)))  (SURVEY §8e)."""
    per = -(-N // G)
    lo = min(N, g * per)
    return lo, min(N, lo + per)


# ----------------------------------------------------------------------------------------------
# tie/rounding-aware comparison used by the parity tests
# ----------------------------------------------------------------------------------------------

def check_topk(idx, val, S_exact: np.ndarray, k: int, rtol: float, atol: float = 0.0):
    """Assert (idx, val) is a valid top-k of the exact score matrix ``S_exact`` ([B, N], float64
    or float32): every returned score matches the exact score of its index within tolerance,
    rows are descending, and any element that differs from the exact top-k set is inside the
    tolerance band around the k-th score (an unavoidable rounding near-tie).  Returns the number
    of such near-tie substitutions."""
    idx = np.asarray(idx)
    val = np.asarray(val, dtype=np.float64)
    B, N = S_exact.shape
    kk = min(k, N)
    swaps = 0
    for b in range(B):
        row = S_exact[b].astype(np.float64)
        got = idx[b, :kk]
        assert (got >= 0).all() and (got < N).all(), f"row {b}: index out of range {got}"
        assert len(set(got.tolist())) == kk, f"row {b}: duplicate indices {got}"
        ref_s = row[got]
        tol = atol + rtol * np.abs(ref_s)
        assert (np.abs(val[b, :kk] - ref_s) <= tol).all(), (
            f"row {b}: score mismatch {val[b, :kk]} vs {ref_s}")
        assert (np.diff(val[b, :kk]) <= 0).all(), f"row {b}: not descending {val[b, :kk]}"
        order = np.argsort(-row, kind="stable")[:kk]
        kth = row[order[-1]]
        band = atol + rtol * abs(kth)
        missing = set(order.tolist()) - set(got.tolist())
        extra = set(got.tolist()) - set(order.tolist())
        for j in missing | extra:
            assert abs(row[j] - kth) <= 2 * band, (
                f"row {b}: index {j} (score {row[j]}) differs from exact top-{kk} "
                f"(k-th score {kth}) by more than rounding")
        swaps += len(extra)
        if kk < k:
            assert (idx[b, kk:] == -1).all()
    return swaps


def check_topk_candidates(idx, val, cand_idx, cand_score, got_exact, k: int, rtol: float, atol: float = 0.0):
    """``check_topk`` for corpora too large to hold the full score matrix.  ``cand_idx`` / ``cand_score``
    ([B, M], M > k) are the exact best M documents per query from an INDEPENDENT computation (ordered score
    desc, index asc); ``got_exact`` [B, k] holds the exact scores of the returned indices.  Same bar: returned
    scores within tolerance, rows descending, and every difference from the exact top-k set lies inside the
    rounding band around the k-th score.  Returns the number of near-tie substitutions."""
    idx = np.asarray(idx)
    val = np.asarray(val, dtype=np.float64)
    cand_idx = np.asarray(cand_idx)
    cand_score = np.asarray(cand_score, dtype=np.float64)
    got_exact = np.asarray(got_exact, dtype=np.float64)
    B, M = cand_idx.shape
    assert M > k
    swaps = 0
    for b in range(B):
        got = idx[b, :k]
        assert (got >= 0).all() and len(set(got.tolist())) == k, f"row {b}: bad indices {got}"
        tol = atol + rtol * np.abs(got_exact[b])
        assert (np.abs(val[b, :k] - got_exact[b]) <= tol).all(), f"row {b}: score mismatch {val[b, :k]} vs {got_exact[b]}"
        assert (np.diff(val[b, :k]) <= 0).all(), f"row {b}: not descending"
        kth = cand_score[b, k - 1]
        band = atol + rtol * abs(kth)
        want = set(cand_idx[b, :k].tolist())
        have = set(got.tolist())
        exact_of = dict(zip(got.tolist(), got_exact[b].tolist()))
        exact_of.update(zip(cand_idx[b].tolist(), cand_score[b].tolist()))
        for j in (want - have) | (have - want):
            assert abs(exact_of[j] - kth) <= 2 * band, (
                f"row {b}: index {j} (score {exact_of[j]}) differs from the exact top-{k} (k-th score {kth}) "
                f"by more than rounding")
        swaps += len(have - want)
    return swaps


# ----------------------------------------------------------------------------------------------
# class-level restatements (return conventions, validation order)
# ----------------------------------------------------------------------------------------------

class SimpleRetrieverPort:
    """NumPy CPU path, raw dot product, one GEMV per query (retriever.py:57-166).  This is the
    code path timed as ``cpu_baseline`` / ``--impl reference`` (kind "port")."""

    def __init__(self, doc_embeddings: np.ndarray, documents: Sequence[str]):
        if doc_embeddings.dtype != np.float32:          # :71-73
            doc_embeddings = doc_embeddings.astype(np.float32)
        self.doc_embeddings = doc_embeddings
        self.documents = documents

    def retrieve_indices(self, query_emb: np.ndarray, k: int) -> np.ndarray:
        N = self.doc_embeddings.shape[0]
        if k <= 0:                                       # :89-91
            return np.empty(0, dtype=np.int64)
        if k > N:                                        # :92-94
            k = N
        if query_emb.dtype != np.float32:                # :98-99
            query_emb = query_emb.astype(np.float32)
        if query_emb.ndim != 1:                          # :100-106
            if query_emb.ndim == 2 and query_emb.shape[0] == 1:
                query_emb = query_emb.flatten()
            else:
                return np.empty(0, dtype=np.int64)
        if query_emb.shape[0] != self.doc_embeddings.shape[1]:   # :108-110
            return np.empty(0, dtype=np.int64)
        sims = self.doc_embeddings @ query_emb           # :112
        top = np.argpartition(sims, -k)[-k:]             # :118
        return top[np.argsort(sims[top])[::-1]]          # :120

    def retrieve(self, query_emb: np.ndarray, k: int = 2) -> List[str]:
        return [self.documents[i] for i in self.retrieve_indices(query_emb, k)]   # :123

    def batch_retrieve(self, query_embeddings: np.ndarray, ks: List[int]) -> List[List[str]]:
        if query_embeddings.ndim != 2:                                   # :146-149
            return [[] for _ in ks]
        if query_embeddings.shape[0] != len(ks):                         # :150-153
            return [[] for _ in ks]
        if query_embeddings.shape[1] != self.doc_embeddings.shape[1]:    # :154-156
            return [[] for _ in ks]
        return [self.retrieve(q, ks[i]) for i, q in enumerate(query_embeddings)]   # :160-164


class CupyRetrieverPort:
    """The live class's logic (retriever.py:564-799) with NumPy standing in for CuPy.  Cosine
    similarity, per-row select at max_k, slice to k_i, string gather, empty-list error
    conventions.  Ties: CuPy's argpartition is a full sort, NumPy's is introselect, so tie
    order here is NOT authoritative; use ``search`` for the tie contract."""

    def __init__(self, doc_embeddings: np.ndarray, documents: Sequence[str]):
        if doc_embeddings.dtype != np.float32:           # :576-578
            doc_embeddings = doc_embeddings.astype(np.float32)
        self.doc_embeddings_cp = np.asarray(doc_embeddings)
        self.documents = documents                       # :589
        self.N_A = self.doc_embeddings_cp.shape[0]       # :590
        self.D = self.doc_embeddings_cp.shape[1]         # :591

    def distance_cosine(self, X, Y, epsilon=1e-8):
        return scores_cosine(X, Y, epsilon)

    def retrieve(self, query_emb, k: int = 2) -> List[str]:
        if not isinstance(query_emb, np.ndarray):        # :629-631
            return []
        if query_emb.ndim == 2 and query_emb.shape[0] == 1:   # :632-634
            query_emb = query_emb.flatten()
        elif query_emb.ndim != 1:                        # :635-637
            return []
        if query_emb.shape[0] != self.D:                 # :638-640
            return []
        actual_k = min(k, self.N_A)                      # :643
        if actual_k <= 0:                                # :644-646
            return []
        if query_emb.dtype != np.float32:                # :650-651
            query_emb = query_emb.astype(np.float32)
        s = self.distance_cosine(query_emb.reshape(1, self.D), self.doc_embeddings_cp)[0]   # :656-657
        cand = np.argpartition(s, self.N_A - actual_k)[-actual_k:]    # :662
        top = cand[np.argsort(-s[cand])]                 # :665-669
        return [self.documents[i] for i in top]          # :675

    def batch_retrieve(self, query_embeddings, ks) -> List[List[str]]:
        batch_size = query_embeddings.shape[0]           # :702 (before the type check)
        if not isinstance(query_embeddings, np.ndarray): # :705-707
            return [[] for _ in ks]
        if query_embeddings.ndim != 2:                   # :708-710
            return [[] for _ in ks]
        if batch_size == 0:                              # :711-713
            return []
        if not isinstance(ks, list) or batch_size != len(ks):   # :714-716
            return [[] for _ in range(batch_size)]
        if query_embeddings.shape[1] != self.D:          # :717-719
            return [[] for _ in ks]
        if not all(isinstance(k_val, int) for k_val in ks):     # :720-722
            return [[] for _ in ks]
        valid_ks = [k_val for k_val in ks if k_val > 0]  # :725
        if not valid_ks:                                 # :726-728
            return [[] for _ in ks]
        max_k = min(max(valid_ks), self.N_A)             # :729-730
        if max_k <= 0:                                   # :731-732
            return [[] for _ in ks]
        if query_embeddings.dtype != np.float32:         # :736-737
            query_embeddings = query_embeddings.astype(np.float32)
        S = self.distance_cosine(query_embeddings, self.doc_embeddings_cp)   # :742
        top = np.empty((batch_size, max_k), dtype=np.int64)                  # :746
        for i in range(batch_size):                                          # :750-769
            cand = np.argpartition(S[i], self.N_A - max_k)[-max_k:]
            top[i] = cand[np.argsort(-S[i][cand])]
        out = []
        for i in range(batch_size):                                          # :776-790
            k_i = ks[i]
            if k_i <= 0:
                out.append([])
                continue
            row = top[i]
            out.append([self.documents[j] for j in row[row >= 0][:k_i]])
        return out


# ----------------------------------------------------------------------------------------------
# synthetic corpus generator (counter-based; any shard of any row range reproduces the same
# rows; the CUDA generator csrc/rs_ingest.cu:synth_fill_kernel is bit-identical)
# ----------------------------------------------------------------------------------------------

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
_GOLD = np.uint64(0x9E3779B97F4A7C15)
_MIX1 = np.uint64(0xBF58476D1CE4E5B9)
_MIX2 = np.uint64(0x94D049BB133111EB)


def _mix64(h: np.ndarray, seed: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (h + np.uint64(1)) * _GOLD
        z = (z ^ (z >> np.uint64(30))) * _MIX1
        z = (z ^ (z >> np.uint64(27))) * _MIX2
        z = z ^ (z >> np.uint64(31))
    return z


def synth_rows(seed: int, row_ids: np.ndarray, D: int = 1024) -> np.ndarray:
    """Rows ``row_ids`` (global indices) of the synthetic unit-normalised corpus, fp32.

    Element 4g+i of row r takes 16 bits of  mix64(seed, r*(D/4)+g);  v = 2*bits - 65535 (odd
    integer), the row is scaled by 1/sqrt(sum v^2) in IEEE fp64 and rounded once to fp32."""
    assert D % 4 == 0
    rows = np.asarray(row_ids, dtype=np.uint64).reshape(-1, 1)
    g = np.arange(D // 4, dtype=np.uint64).reshape(1, -1)
    with np.errstate(over="ignore"):
        z = _mix64(rows * np.uint64(D // 4) + g, seed)
    v = np.empty((rows.shape[0], D), dtype=np.int64)
    for i in range(4):
        v[:, i::4] = ((z >> np.uint64(16 * i)) & np.uint64(0xFFFF)).astype(np.int64) * 2 - 65535
    sumsq = (v * v).sum(axis=1, keepdims=True)
    rinv = 1.0 / np.sqrt(sumsq.astype(np.float64))
    return (v.astype(np.float64) * rinv).astype(np.float32)


def synth_queries(seed: int, n_rows: int, B: int, D: int = 1024, noise: float = 0.01,
                  corpus_seed: int = 1234, dtype: int = DTYPE_FP32):
    """B queries = seeded corpus rows + noise*N(0,1)/sqrt(D)-scaled perturbation, re-normalised
    and rounded to be exactly representable in ``dtype``.  Returns
    (queries fp32 [B, D], source row ids)."""
    rng = np.random.default_rng(seed)
    ids = rng.integers(0, n_rows, size=B, dtype=np.int64)
    base = synth_rows(corpus_seed, ids, D).astype(np.float64)
    q = base + noise * rng.standard_normal((B, D))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return round_to_dtype(q.astype(np.float32), dtype), ids


# ----------------------------------------------------------------------------------------------
# HBM tile layout (csrc/rs_layout.h): the corpus is stored as UMMA-canonical K-major
# SWIZZLE_128B blocks so that one contiguous bulk copy lands an MMA-ready operand in smem
# ----------------------------------------------------------------------------------------------

TILE_ROWS = 128          # documents per tile (= MMA N)
KCHUNK_BYTES = 128       # one swizzle row: 64 two-byte elements
BLOCK_BYTES = TILE_ROWS * KCHUNK_BYTES


def tile_layout_offsets(n_rows: int, D: int, planes: int = 1) -> np.ndarray:
    """Byte offset (into the tiled buffer) of every two-byte element [plane, row, col]."""
    assert D % 64 == 0
    KC = D // 64
    r = np.arange(n_rows, dtype=np.int64).reshape(1, -1, 1)
    c = np.arange(D, dtype=np.int64).reshape(1, 1, -1)
    p = np.arange(planes, dtype=np.int64).reshape(-1, 1, 1)
    t, rr = r // TILE_ROWS, r % TILE_ROWS
    kc, cc = c // 64, c % 64
    chunk, within = cc // 8, cc % 8
    tile_bytes = planes * KC * BLOCK_BYTES
    return (t * tile_bytes + (p * KC + kc) * BLOCK_BYTES + (rr // 8) * 1024 + (rr % 8) * 128
            + ((chunk ^ (rr % 8)) * 16) + within * 2)


def tile_corpus(planes_u16: np.ndarray) -> np.ndarray:
    """[planes, N, D] uint16 bit patterns -> tiled byte image (uint16 array), rows past N zero."""
    P, N, D = planes_u16.shape
    n_tiles = -(-N // TILE_ROWS)
    out = np.zeros(n_tiles * P * (D // 64) * BLOCK_BYTES // 2, dtype=np.uint16)
    off = tile_layout_offsets(N, D, P) // 2
    out[off.reshape(-1)] = planes_u16.reshape(-1)
    return out


def corpus_planes_u16(D_f32: np.ndarray, dtype: int, scale_exp: int = 0) -> np.ndarray:
    """The 16-bit planes stored on device for an fp32 input corpus in storage ``dtype``."""
    if dtype == DTYPE_BF16:
        return bf16_bits(D_f32)[None]
    if dtype == DTYPE_FP16:
        return np.asarray(D_f32, dtype=np.float32).astype(np.float16).view(np.uint16)[None]
    hi, lo = split_fp16(D_f32, scale_exp)
    return np.stack([hi.view(np.uint16), lo.view(np.uint16)])


def stored_values(D_f32: np.ndarray, dtype: int, scale_exp: int = 0) -> np.ndarray:
    """The exact real values the device corpus represents (float64)."""
    if dtype == DTYPE_FP32:
        hi, lo = split_fp16(D_f32, scale_exp)
        return unsplit_fp16(hi, lo, scale_exp)
    return round_to_dtype(D_f32, dtype).astype(np.float64)


logging.getLogger(__name__).addHandler(logging.NullHandler())
