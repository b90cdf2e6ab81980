"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against
the oracle on the same seeded inputs and against the reference-generated golden fixtures.

Bars (north_star): top-k index sets identical (ties -> lower index; a substitution is accepted
only inside the rounding band around the k-th score and is reported), scores within 1e-5
relative for fp32 storage, 2e-3 for bf16/fp16."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import retrieval_oracle as o
from rag_serving_system_b200 import CupyRetriever, SimpleRetriever, B200Retriever, _cabi

pytestmark = pytest.mark.gpu

DT = {"bf16": o.DTYPE_BF16, "fp16": o.DTYPE_FP16, "fp32": o.DTYPE_FP32}
RTOL = {"bf16": 2e-3, "fp16": 2e-3, "fp32": 1e-5}


def make_index(E, dtype, scale_exp=0, **kw):
    ix = _cabi.Index(E.shape[0], E.shape[1], dtype, scale_exp=scale_exp, **kw)
    ix.load_rows(E)
    ix.finalize()
    return ix


def exact_scores(Q, E, dtype, metric, terms, scale_exp=0):
    Dv = o.stored_values(E, DT[dtype], scale_exp)
    Qv = Q.astype(np.float64) if (terms == 2 or dtype == "fp32") else o.round_to_dtype(Q, DT[dtype]).astype(np.float64)
    return o.scores(Qv, Dv, metric, accum=np.float64)


def run_case(E, Q, k, dtype, metric, terms, scale_exp=0, exact_idx=False):
    ix = make_index(E, dtype, scale_exp, max_batch=max(64, len(Q)), max_k=max(16, k))
    try:
        idx, sc = ix.search(Q, k, metric, terms)
        S = exact_scores(Q, E, dtype, metric, terms, scale_exp)
        swaps = o.check_topk(idx, sc, S, k, rtol=RTOL[dtype], atol=1e-6)
        if exact_idx:
            eidx, _ = o.topk_rows(S, k)
            assert (idx == eidx).all()
        return idx, sc, swaps
    finally:
        ix.close()


# ---------------------------------------------------------------- ingest / layout: bit-exact
@pytest.mark.parametrize("dtype", ["bf16", "fp16", "fp32"])
def test_ingest_image_and_roundtrip_bit_exact(squad, dtype):
    E = squad.E[:333]                                     # ragged: not a multiple of the 128-row tile
    se = 4 if dtype == "fp32" else 0
    ix = make_index(E, dtype, se)
    try:
        img = ix.get_image(0, ix.stats()["corpus_bytes"]).view(np.uint16)
        assert np.array_equal(img, o.tile_corpus(o.corpus_planes_u16(E, DT[dtype], se)))
        rows = ix.get_rows(0, len(E))
        assert np.array_equal(rows, o.stored_values(E, DT[dtype], se).astype(np.float32))
        nr = ix.get_norms(0, len(E))
        assert np.allclose(nr, np.linalg.norm(rows.astype(np.float64), axis=1), rtol=1e-6)
        part = ix.get_rows(130, 70)
        assert np.array_equal(part, rows[130:200])
    finally:
        ix.close()


@pytest.mark.parametrize("dtype", ["bf16", "fp16", "fp32"])
def test_synthetic_generator_bit_exact_any_shard(dtype):
    N, D, g0 = 700, 1024, 123_456_789
    ix = _cabi.Index(N, D, dtype, row_offset=g0, scale_exp=4 if dtype == "fp32" else 0)
    try:
        ix.fill_synthetic(1234)                           # local row i == global row g0 + i
        ix.finalize()
        exp32 = o.synth_rows(1234, g0 + np.arange(N), D)
        se = 4 if dtype == "fp32" else 0
        assert np.array_equal(ix.get_rows(0, N), o.stored_values(exp32, DT[dtype], se).astype(np.float32))
    finally:
        ix.close()


def test_bf16_source_rows_load_bit_exact(squad):
    bits = o.bf16_bits(squad.E[:200])
    ix = _cabi.Index(200, 1024, "bf16")
    try:
        ix.load_rows(bits, 0, src_dtype="bf16")
        ix.load_rows(squad.E[100:200], 100)               # chunked ingest at an offset, fp32 source
        ix.finalize()
        assert np.array_equal(ix.get_rows(0, 200), o.round_to_bf16(squad.E[:200]))
    finally:
        ix.close()


def test_from_npy_memmap_sources(tmp_path, squad):
    """.npy corpora are memory-mapped and ingested chunk by chunk; fp32 / fp16 files
    load directly, bf16 arrives as a uint16 file (+ source_dtype) or as ml_dtypes' '<V2'."""
    E, Q, k = squad.E, squad.Q, 10
    np.save(tmp_path / "f32.npy", E)
    np.save(tmp_path / "f16.npy", E.astype(np.float16))
    np.save(tmp_path / "bf16_u16.npy", o.bf16_bits(E))
    cases = [("f32.npy", {"dtype": "fp32"}, E), ("f16.npy", {"dtype": "fp16"}, E.astype(np.float16).astype(np.float32)),
             ("bf16_u16.npy", {"dtype": "bf16", "source_dtype": "bf16"}, o.round_to_bf16(E))]
    try:
        import ml_dtypes
        np.save(tmp_path / "bf16_v2.npy", E.astype(ml_dtypes.bfloat16))
        cases.append(("bf16_v2.npy", {"dtype": "bf16"}, o.round_to_bf16(E)))
    except ImportError:
        pass
    for name, kw, stored in cases:
        r = B200Retriever.from_npy(str(tmp_path / name), metric="dot", **kw)
        assert r.N_A == len(E) and r.documents[3] == "doc-3"
        idx, sc = r.search(Q, k)
        S = o.scores_dot(Q.astype(np.float64), stored.astype(np.float64), np.float64)
        o.check_topk(idx, sc, S, k, rtol=RTOL[kw["dtype"]], atol=1e-6)
        r.close()


def test_device_tensor_corpus_and_query_handoff(squad):
    """corpus and queries that already live on the GPU never touch the host."""
    import torch
    E, Q, k = squad.E, squad.Q, 10
    for tdt, name in ((torch.float32, "fp32"), (torch.bfloat16, "bf16"), (torch.float16, "fp16")):
        emb = torch.from_numpy(E).cuda().to(tdt)
        r = B200Retriever.from_device_tensor(emb, squad.contexts, metric="dot")
        stored = emb.float().cpu().numpy()
        qd = torch.from_numpy(Q).cuda()
        oi, os_ = r.search_tensor(qd, k)
        torch.cuda.synchronize()
        idx, sc = oi.cpu().numpy(), os_.cpu().numpy()
        terms = 2 if name == "fp32" else 1
        Qv = Q if terms == 2 else o.round_to_dtype(Q, DT[name])
        S = o.scores_dot(Qv.astype(np.float64), stored.astype(np.float64), np.float64)
        o.check_topk(idx, sc, S, k, rtol=RTOL[name], atol=1e-6)
        hidx, hsc = r.search(Q, k, query_terms=terms)
        assert np.array_equal(hidx, idx) and np.array_equal(hsc, sc)
        r.close()
    # pinned host source: DMA straight from the caller's buffer
    pin = torch.from_numpy(E).pin_memory()
    ix = _cabi.Index(len(E), 1024, "bf16")
    try:
        ix.load_rows_ptr(pin.data_ptr(), len(E), "fp32")
        ix.finalize()
        assert np.array_equal(ix.get_rows(0, len(E)), o.round_to_bf16(E))
    finally:
        ix.close()


def test_large_ingest_double_buffered():
    """Several staging chunks (> 32 MiB each way) through the double-buffered pageable path."""
    N, D = 40_000, 1024
    E = o.synth_rows(3, np.arange(N), D)
    ix = _cabi.Index(N, D, "fp32", scale_exp=4)
    try:
        ix.load_rows(E)                                    # 160 MB in one call: 5 chunks
        ix.finalize()
        got = np.concatenate([ix.get_rows(r, min(10_000, N - r)) for r in range(0, N, 10_000)])
        assert np.array_equal(got, o.stored_values(E, o.DTYPE_FP32, 4).astype(np.float32))
    finally:
        ix.close()


# ---------------------------------------------------------------- golden fixtures (reference outputs)
def test_fp32_dot_equals_reference_golden(golden):
    idx, sc, _ = run_case(golden.E, golden.Q, golden.k, "fp32", "dot", 2, scale_exp=4)
    assert (idx == golden.idx_fixed).all()               # identical to the unmodified SimpleRetriever


def test_drop_in_classes_equal_reference_golden(golden):
    s = SimpleRetriever(golden.E, golden.contexts)
    assert s.batch_retrieve(golden.Q, [golden.k] * len(golden.Q)) == golden.texts
    got = s.batch_retrieve(golden.Q, golden.ks_var)
    for row, exp in zip(got, golden.idx_var):
        assert row == [golden.contexts[i] for i in exp if i >= 0]
    assert s.retrieve(golden.E[7], 1) == [golden.contexts[7]]
    s.close()
    c = CupyRetriever(golden.E, golden.contexts)          # cosine: same ranking on unit-norm corpora
    assert c.batch_retrieve(golden.Q, [golden.k] * len(golden.Q)) == golden.texts
    ref = o.CupyRetrieverPort(golden.E, golden.contexts)
    assert c.batch_retrieve(golden.Q, golden.ks_var) == ref.batch_retrieve(golden.Q, golden.ks_var)
    assert c.retrieve(golden.Q[1], 3) == ref.retrieve(golden.Q[1], 3)
    assert c.batch_retrieve(golden.Q[:, :5], [1] * len(golden.Q)) == [[] for _ in golden.Q]
    idx, sc, ctx = c.retrieve_with_scores(golden.Q, golden.k)
    S = o.scores_cosine(golden.Q.astype(np.float64), golden.E.astype(np.float64), accum=np.float64)
    o.check_topk(idx, sc, S, golden.k, rtol=1e-5, atol=1e-7)
    assert ctx == golden.texts
    c.close()


def test_cosine_equals_reference_cupy_class_golden(golden):
    # the unmodified reference CupyRetriever (NumPy standing in for CuPy): indices identical,
    # scores within 1e-5 relative, strings identical, variable k identical
    Q, k = golden.cos_Q, golden.k
    idx, sc, _ = run_case(golden.E, Q, k, "fp32", "cosine", 2, scale_exp=4)
    assert (idx == golden.cos_idx_fixed).all()
    assert np.allclose(sc, golden.cos_scores_fixed, rtol=1e-5, atol=0)
    c = CupyRetriever(golden.E, golden.contexts)
    assert c.batch_retrieve(Q, [k] * len(Q)) == golden.cos_texts
    got = c.batch_retrieve(Q, golden.ks_var)
    for row, exp in zip(got, golden.cos_idx_var):
        assert row == [golden.contexts[i] for i in exp if i >= 0]
    assert c.retrieve(Q[3], k) == golden.cos_texts[3]
    # contexts gathered and joined from the byte arena (rs_docs_join): what batch_processor.py:100 builds per request
    assert c.batch_retrieve_contexts(Q, [k] * len(Q)) == ["\n---\n".join(t) for t in golden.cos_texts]
    c.close()


def test_self_retrieval_kat_all_dtypes(golden):
    for dtype in ("bf16", "fp16", "fp32"):
        r = B200Retriever(golden.E, golden.contexts, dtype=dtype, metric="cosine")
        idx, sc = r.search(golden.E[:64], 1)
        assert (idx[:, 0] == np.arange(min(64, len(golden.E)))).all()
        assert np.abs(sc[:, 0] - 1.0).max() < (3e-3 if dtype != "fp32" else 1e-5)
        r.close()


# ---------------------------------------------------------------- sweeps against the oracle
@pytest.mark.parametrize("dtype,terms", [("bf16", 1), ("bf16", 2), ("fp16", 1), ("fp16", 2), ("fp32", 2)])
@pytest.mark.parametrize("metric", ["dot", "cosine"])
def test_squad_parity(squad, dtype, terms, metric):
    run_case(squad.E, squad.Q, 10, dtype, metric, terms, scale_exp=4 if dtype == "fp32" else 0)


@pytest.mark.parametrize("B", [1, 2, 7, 8, 31, 33, 64, 65, 100, 200])
def test_batch_sizes(squad, B):
    rng = np.random.default_rng(B)
    Q = (squad.E[rng.integers(0, 1000, B)] + 0.05 * rng.standard_normal((B, 1024))).astype(np.float32)
    run_case(squad.E, Q, 5, "bf16", "dot", 1)
    if B in (1, 33, 100):
        run_case(squad.E, Q, 5, "fp32", "cosine", 2, scale_exp=4)


@pytest.mark.parametrize("k", [1, 2, 5, 10, 16, 17, 32, 100, 128])
def test_k_values(squad, k):
    run_case(squad.E, squad.Q[:16], k, "bf16", "dot", 1)
    if k in (1, 16, 17, 100):
        run_case(squad.E, squad.Q[:16], k, "fp32", "cosine", 2, scale_exp=4)


@pytest.mark.parametrize("n", [1, 5, 127, 128, 129, 255, 1000])
def test_ragged_and_tiny_corpora(squad, n):
    idx, sc, _ = run_case(squad.E[:n], squad.Q[:8], 10, "bf16", "dot", 1)
    if n < 10:                                            # N < k: -1 / -inf padding (retriever.py:767-769)
        assert (idx[:, n:] == -1).all() and np.isneginf(sc[:, n:]).all()


@pytest.mark.parametrize("D", [64, 128, 384, 768, 2048])
def test_other_dims(D):
    rng = np.random.default_rng(D)
    E = rng.standard_normal((600, D)).astype(np.float32)
    Q = rng.standard_normal((12, D)).astype(np.float32)   # un-normalised, like the service's queries
    run_case(E, Q, 10, "bf16", "cosine", 2)
    run_case(E, Q, 10, "fp32", "dot", 2, scale_exp=retriever_scale(E))


def retriever_scale(E):
    from rag_serving_system_b200.retriever import _fp32_scale_exp
    return _fp32_scale_exp(float(np.abs(E).max()))


def test_many_tiles_per_cta_and_all_ctas():
    N = 128 * 148 * 3 + 77
    E = o.synth_rows(7, np.arange(N))
    rng = np.random.default_rng(1)
    src = rng.integers(0, N, 32)
    src[:4] = [0, N - 1, 128 * 148, 128 * 148 * 3]        # first row, last (ragged) row, tile boundaries
    Q = o.round_to_bf16(E[src] + 0.02 * rng.standard_normal((32, 1024)).astype(np.float32))
    for dtype, terms, metric in (("bf16", 1, "dot"), ("fp16", 1, "cosine"), ("fp32", 2, "dot")):
        idx, sc, _ = run_case(E, Q, 10, dtype, metric, terms, scale_exp=4 if dtype == "fp32" else 0)
        assert (idx[:, 0] == src).all()


# ---------------------------------------------------------------- every kernel mode the planner can pick
MODES = [
    {"RS_MMA_M": "128"}, {"RS_MMA_M": "64"}, {"RS_CPS": "1"}, {"RS_CPS": "1", "RS_QRESIDENT": "0", "RS_SUB": "2", "RS_NBUF": "2"},
    {"RS_QRESIDENT": "0", "RS_SUB": "1", "RS_NBUF": "2"}, {"RS_QRESIDENT": "0", "RS_SUB": "2", "RS_NBUF": "2"},
    {"RS_QRESIDENT": "0", "RS_SUB": "4", "RS_NBUF": "1"}, {"RS_QRESIDENT": "0", "RS_SUB": "2", "RS_NBUF": "1", "RS_MMA_M": "128"},
    {"RS_QRESIDENT": "1", "RS_SUB": "2", "RS_NBUF": "2"},
    # the multi-launch fallbacks of the single-launch search: separate query_prep_kernel / merge kernel
    {"RS_INLINE_PREP": "0"}, {"RS_FUSED_TAIL": "0"}, {"RS_INLINE_PREP": "0", "RS_FUSED_TAIL": "0"},
    # k > 16 state: lists in shared memory instead of global memory, no bound shared between CTAs
    {"RS_LISTS_GLOBAL": "0"}, {"RS_SHARED_BOUND": "0"}, {"RS_DYN_SCHED": "0"},
    # query-major kernel everywhere / doc-major kernel only for one-term 16-bit searches
    {"RS_DOCM": "0"}, {"RS_DOCM": "2"}, {"RS_DOCM": "0", "RS_FUSED_TAIL": "0", "RS_INLINE_PREP": "0"},
]


@pytest.mark.parametrize("mode", MODES, ids=lambda m: ",".join(f"{k[3:]}={v}" for k, v in m.items()))
def test_kernel_modes(monkeypatch, squad, mode):
    for k_, v in mode.items():
        monkeypatch.setenv(k_, v)
    N = 128 * 148 * 2 + 300                                # supertiles that end in partial groups
    E = o.synth_rows(11, np.arange(N))
    rng = np.random.default_rng(2)
    src = rng.integers(0, N, 24)
    src[:2] = [0, N - 1]
    Q = o.round_to_bf16(E[src] + 0.02 * rng.standard_normal((24, 1024)).astype(np.float32))
    idx, _, _ = run_case(E, Q, 10, "bf16", "dot", 1)
    assert (idx[:, 0] == src).all()
    run_case(squad.E, squad.Q, 10, "bf16", "cosine", 1)
    run_case(squad.E[:700], squad.Q[:9], 20, "bf16", "dot", 1)           # shared-memory list path
    if mode.get("RS_SUB") != "4":                                       # two accumulators: sub*nbuf*2 <= 4
        nb = {"RS_NBUF": "1"} if mode.get("RS_SUB") == "2" else {}
        for k_, v in nb.items():
            monkeypatch.setenv(k_, v)
        run_case(squad.E, squad.Q, 10, "fp32", "dot", 2, scale_exp=4)
        run_case(E[:40_000], Q, 10, "fp16", "cosine", 2)


def test_single_launch_search_counts_and_small_grids(squad):
    """One fused launch per search by default; extra merge-only CTAs when the corpus has fewer tiles than
    the batch has queries (grid >= B so every query owns a tail CTA)."""
    for n, B in ((100, 32), (128, 64), (300, 17), (1000, 32), (1000, 5)):
        ix = make_index(squad.E[:n], "bf16", max_batch=64, max_k=16)
        try:
            Q = squad.Q[np.arange(B) % len(squad.Q)] * (1 + 0.01 * np.arange(B)[:, None]).astype(np.float32)
            for _ in range(3):                               # repeated launches reuse the reset counters
                idx, sc = ix.search(Q, 10, "dot", 1)
            st = ix.stats()
            assert st["launches"] == (1 if B <= 32 else 2), st      # B > 32: query_prep_kernel stays separate
            assert st["grid"] == max(min(148, -(-n // 128)), min(B, 148))
            S = exact_scores(Q, squad.E[:n], "bf16", "dot", 1)
            o.check_topk(idx, sc, S, 10, rtol=2e-3, atol=1e-6)
        finally:
            ix.close()


@pytest.mark.parametrize("B,dtype,terms", [(96, "bf16", 1), (128, "bf16", 1), (128, "fp32", 2), (128, "fp16", 2),
                                           (65, "bf16", 2), (256, "bf16", 1)])
def test_large_batches_stream_queries_single_pass(squad, B, dtype, terms):
    rng = np.random.default_rng(B)
    N = 128 * 148 + 50
    E = o.synth_rows(5, np.arange(N))
    Q = (E[rng.integers(0, N, B)] + 0.05 * rng.standard_normal((B, 1024))).astype(np.float32)
    ix = make_index(E, dtype, 4 if dtype == "fp32" else 0, max_batch=256, max_k=16)
    try:
        idx, sc = ix.search(Q, 10, "dot", terms)
        st = ix.stats()
        assert st["passes"] == 1 and st["q_resident"] == 0      # B = 256: the wide-batch kernel, still one pass
        S = exact_scores(Q, E, dtype, "dot", terms, 4 if dtype == "fp32" else 0)
        o.check_topk(idx, sc, S, 10, rtol=RTOL[dtype], atol=1e-6)
    finally:
        ix.close()


@pytest.mark.parametrize("k", [17, 33, 64, 100, 128])
def test_large_k_buffered_selection(k):
    """k > 16: sorted list + candidate buffer in shared memory, warp-cooperative merges."""
    N = 128 * 148 * 2 + 99
    E = o.synth_rows(21, np.arange(N))
    E[5000:5040] = E[17]                                   # 41 identical rows: a long tie run inside the top-k
    rng = np.random.default_rng(k)
    src = rng.integers(0, N, 40)
    src[0] = 17
    Q = o.round_to_bf16(E[src] + 0.05 * rng.standard_normal((40, 1024)).astype(np.float32))
    Q[0] = o.round_to_bf16(E[17])
    idx, sc, _ = run_case(E, Q, k, "bf16", "dot", 1)
    tie = [17] + list(range(5000, 5040))
    assert idx[0, :min(k, 41)].tolist() == tie[:min(k, 41)]  # ties in ascending index order
    run_case(E[:30_000], Q[:9], k, "fp32", "cosine", 2, scale_exp=4)
    run_case(E[:k - 3], Q[:5], k, "bf16", "dot", 1)          # N < k
    B = 128
    Qb = o.round_to_bf16(E[rng.integers(0, N, B)] + 0.05 * rng.standard_normal((B, 1024)).astype(np.float32))
    run_case(E[:50_000], Qb, k, "bf16", "dot", 1)            # streamed queries + shared-memory lists


# ---------------------------------------------------------------- ties, degenerate inputs
def test_duplicate_rows_tie_to_lower_index(squad):
    E = squad.E[:600].copy()
    E[100] = E[0]
    E[599] = E[0]
    E[300] = E[200]
    Q = np.stack([E[0], E[200]])
    for dtype, terms in (("bf16", 1), ("fp32", 2)):
        for metric in ("dot", "cosine"):
            idx, sc, _ = run_case(E, Q, 4, dtype, metric, terms, scale_exp=4 if dtype == "fp32" else 0)
            assert idx[0, :3].tolist() == [0, 100, 599] and idx[1, :2].tolist() == [200, 300]
            i1, _, _ = run_case(E, Q, 1, dtype, metric, terms, scale_exp=4 if dtype == "fp32" else 0)
            assert i1[:, 0].tolist() == [0, 200]          # at the k boundary the LOWER index stays


def test_all_equal_rows_and_zero_query():
    E = np.tile(np.linspace(-1, 1, 256, dtype=np.float32), (700, 1))
    Q = np.stack([E[0], -E[0], np.zeros(256, np.float32)])
    for metric in ("dot", "cosine"):
        idx, sc, _ = run_case(E, Q, 12, "bf16", metric, 1, exact_idx=True)
        assert (idx == np.arange(12)).all()               # every score ties -> first k rows, in order


def test_shared_bound_keeps_ties_at_the_bound():
    """k > 16 with at least k CTAs at work: the bound shared between CTAs (min over CTAs of each CTA's best
    score) must keep documents that TIE with it.  All rows identical -> every score equals the bound and the
    answer is the first k rows; a zero query makes the bound exactly 0; a corpus whose best rows sit in one
    tile makes one CTA hold the whole top-k while every other CTA only publishes a low best."""
    N = 128 * 148 * 2 + 50
    E = np.tile(np.linspace(-1, 1, 256, dtype=np.float32), (N, 1))
    Q = np.stack([E[0], -E[0], np.zeros(256, np.float32)])
    for k in (17, 100):
        for metric in ("dot", "cosine"):
            idx, sc, _ = run_case(E, Q, k, "bf16", metric, 1, exact_idx=True)
            assert (idx == np.arange(k)).all()
    # the same through the wide-batch (CTA-pair) kernel, whose sample pass applies the bound before anything is appended:
    # a zero query's cosine threshold sits just below 0 and its raw-score bound must not underflow to -0
    Qw = np.concatenate([Q, np.tile(E[0], (150, 1)) * np.linspace(0.5, 2.0, 150, dtype=np.float32)[:, None]])
    for metric in ("dot", "cosine"):
        idx, sc, _ = run_case(E, Qw, 17, "bf16", metric, 1, exact_idx=True)
        assert (idx == np.arange(17)).all()
    rng = np.random.default_rng(3)
    E2 = o.synth_rows(33, np.arange(N), 256)
    hot = 128 * 177                                            # one tile holds 120 near-copies of the query
    q = o.round_to_bf16(E2[5])
    E2[hot:hot + 120] = q * (1.0 - 1e-3 * np.arange(120, dtype=np.float32))[:, None]
    idx, sc, _ = run_case(E2, np.stack([q, o.round_to_bf16(E2[9])]), 100, "bf16", "dot", 1, exact_idx=True)
    assert set(idx[0, :99].tolist()) <= set(range(hot, hot + 120)) | {5}


def test_negative_scores_and_cosine_clip(squad):
    E = squad.E[:500]
    Q = -squad.Q[:8]                                      # every similarity negative (thr < 0 branch)
    run_case(E, Q, 10, "bf16", "cosine", 1)
    run_case(E, Q * 37.5, 10, "fp32", "cosine", 2, scale_exp=4)
    big = (squad.E[:500] * 50).astype(np.float32)         # un-normalised corpus: norms vary per row
    big[::3] *= 0.01
    run_case(big, squad.Q[:8], 10, "bf16", "cosine", 2)
    run_case(big, squad.Q[:8], 10, "fp32", "dot", 2, scale_exp=retriever_scale(big))


# ---------------------------------------------------------------- merge + multi-shard
def test_merge_device_matches_oracle():
    import torch
    rng = np.random.default_rng(3)
    for G, B, kc, k in ((2, 32, 10, 10), (8, 5, 7, 7), (148, 3, 10, 10), (4, 2, 100, 100), (3, 4, 6, 4)):
        s = np.sort(rng.integers(-5, 6, size=(G, B, kc)).astype(np.float32), axis=2)[:, :, ::-1].copy()
        i = rng.permutation(G * B * kc).reshape(G, B, kc).astype(np.int64)
        i[-1, :, kc // 2:] = -1                            # padding entries ...
        s[i < 0] = -np.inf                                 # ... carry -inf and sit at the END of a list
        order = np.lexsort((i, -s, i < 0), axis=2)         # each list sorted (score desc, idx asc), pads last
        s, i = np.take_along_axis(s, order, 2), np.take_along_axis(i, order, 2)
        ts, ti = torch.from_numpy(s).cuda(), torch.from_numpy(i).cuda()
        oi = torch.empty((B, k), dtype=torch.int64, device="cuda")
        os_ = torch.empty((B, k), dtype=torch.float32, device="cuda")
        _cabi.merge_device(0, ts.data_ptr(), ti.data_ptr(), G, B, kc, k, oi.data_ptr(), os_.data_ptr(),
                           torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        ei, es = o.merge_candidates(s, i, k)
        assert np.array_equal(oi.cpu().numpy(), ei) and np.array_equal(os_.cpu().numpy(), es)


def test_row_offset_and_two_shards_equal_one(squad):
    E, Q, k = squad.E, squad.Q, 10
    full, _, _ = run_case(E, Q, k, "bf16", "dot", 1)
    a = _cabi.Index(600, 1024, "bf16", row_offset=0)
    b = _cabi.Index(400, 1024, "bf16", row_offset=600)
    try:
        a.load_rows(E[:600]); a.finalize()
        b.load_rows(E[600:]); b.finalize()
        ia, sa = a.search(Q, k, "dot", 1)
        ib, sb = b.search(Q, k, "dot", 1)
        assert ib.min() >= 600
        mi, ms = o.merge_candidates(np.stack([sa, sb]), np.stack([ia, ib]), k)
        assert (mi == full).all()
    finally:
        a.close(); b.close()


def test_device_buffer_entry_point(squad):
    import torch
    ix = make_index(squad.E, "bf16")
    try:
        q = torch.from_numpy(squad.Q).cuda()
        oi = torch.empty((32, 10), dtype=torch.int64, device="cuda")
        os_ = torch.empty((32, 10), dtype=torch.float32, device="cuda")
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            ix.search_device(q.data_ptr(), 32, 10, oi.data_ptr(), os_.data_ptr(), "dot", 1, st.cuda_stream)
        st.synchronize()
        hi, hs = ix.search(squad.Q, 10, "dot", 1)
        assert np.array_equal(oi.cpu().numpy(), hi) and np.array_equal(os_.cpu().numpy(), hs)
        assert ix.kernel_times_ms(2).min() > 0
    finally:
        ix.close()


def test_graph_replay_uses_fresh_queries(squad, monkeypatch):
    """Host-buffer searches replay a captured CUDA graph from the third call of a shape on; every
    replay must see that call's queries and a changed shape must not reuse the wrong graph."""
    ix = make_index(squad.E, "bf16")
    try:
        rng = np.random.default_rng(4)
        for it in range(6):
            B, k = (32, 10) if it != 3 else (7, 5)
            Q = o.round_to_bf16(squad.E[rng.integers(0, 1000, B)] + 0.05 * rng.standard_normal((B, 1024)).astype(np.float32))
            idx, sc = ix.search(Q, k, "dot", 1)
            S = o.scores_dot(Q.astype(np.float64), o.round_to_bf16(squad.E).astype(np.float64), np.float64)
            o.check_topk(idx, sc, S, k, rtol=2e-3, atol=1e-6)
            eidx, _ = o.topk_rows(S, k)
            assert (idx[:, 0] == eidx[:, 0]).all()
        ix.set_option("RS_GRAPH", 0)                        # (the environment is only read by rs_create)
        i2, s2 = ix.search(Q, k, "dot", 1)                  # eager path gives the same answer
        assert np.array_equal(i2, idx) and np.array_equal(s2, sc)
    finally:
        ix.close()


def test_errors_are_loud(squad):
    ix = _cabi.Index(100, 1024, "bf16", max_batch=8, max_k=4)
    try:
        with pytest.raises(_cabi.RsError, match="rs_finalize"):
            ix.search(squad.Q[:2], 2)
        ix.load_rows(squad.E[:100]); ix.finalize()
        with pytest.raises(_cabi.RsError):
            ix.search(squad.Q[:9], 2)                      # B > max_batch
        with pytest.raises(_cabi.RsError):
            ix.search(squad.Q[:2], 5)                      # k > max_k
        with pytest.raises(_cabi.RsError):
            ix.load_rows(squad.E[:10], 95)                 # rows out of range
    finally:
        ix.close()


# ---------------------------------------------------------------- BASELINE sizes: size-independent properties
def test_config3_1m_fp32_torch_crosscheck():
    """1M x 1024 fp32 (BASELINE configs[2]): cross-check against a torch fp32 restatement
    (TF32 off) of the same arithmetic - labelled a restatement, not the reference."""
    import torch
    torch.backends.cuda.matmul.allow_tf32 = False
    N, D, k = 1_000_000, 1024, 10
    ix = _cabi.Index(N, D, "fp32", scale_exp=4, max_batch=128, max_k=16)
    try:
        ix.fill_synthetic(1234, 0)
        ix.finalize()
        Q, src = o.synth_queries(11, N, 128, D, dtype=o.DTYPE_FP32)
        rows = torch.empty((N, D), dtype=torch.float32, device="cuda")
        step = 1 << 15
        for r in range(0, N, step):
            rows[r:r + step] = torch.from_numpy(ix.get_rows(r, min(step, N - r))).cuda()
        for B in (1, 32, 128):
            idx, sc = ix.search(Q[:B], k, "dot", 2)
            assert (idx[:, 0] == src[:B]).all()
            S = (torch.from_numpy(Q[:B]).cuda().double() @ rows.double().T).cpu().numpy()
            o.check_topk(idx, sc, S, k, rtol=1e-5, atol=1e-7)
    finally:
        ix.close()


def torch_exact_topk(ix, Q, metric, m, chunk=1 << 20):
    """INDEPENDENT computation on the GPU (a torch restatement of retriever.py:600-609 + a full-row top-m, not
    the kernel under test): the index's rows are read back untiled (rs_get_rows_device), scored against the
    queries with an fp64 matmul chunk by chunk, and the best m (score desc, index asc) are kept.  Returns
    (cand_idx [B,m], cand_score [B,m] float64, exact(idx [B,k]) -> float64 scores of arbitrary rows)."""
    import torch
    dev = torch.device("cuda", ix.device)
    N, D = ix.n_rows, ix.dim
    q = torch.from_numpy(np.ascontiguousarray(Q, dtype=np.float32)).to(dev).double()
    nq = q.float().norm(dim=1, keepdim=True).double() if metric == "cosine" else None
    buf = torch.empty((min(chunk, N), D), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream

    def score_rows(rows32):
        s_ = q @ rows32.double().T
        if metric == "cosine":                                                   # retriever.py:602-609
            nd = rows32.norm(dim=1, keepdim=True).double()
            s_ = (s_ / (nq @ nd.T + 1e-8)).clamp_(-1.0, 1.0)
        return s_

    best_s = torch.full((q.shape[0], 0), -float("inf"), dtype=torch.float64, device=dev)
    best_i = torch.zeros((q.shape[0], 0), dtype=torch.int64, device=dev)
    for r in range(0, N, chunk):
        n = min(chunk, N - r)
        ix.get_rows_device(r, n, buf.data_ptr(), stream)
        s_ = score_rows(buf[:n])
        ts, ti = torch.topk(s_, min(m, n), dim=1)
        best_s = torch.cat([best_s, ts], dim=1)
        best_i = torch.cat([best_i, ti + r], dim=1)
        if best_s.shape[1] > 4 * m:
            ts, sel = torch.topk(best_s, m, dim=1)
            best_s, best_i = ts, torch.gather(best_i, 1, sel)
    cs, ci = best_s.cpu().numpy(), best_i.cpu().numpy()
    order = np.stack([np.lexsort((ci[b], -cs[b]))[:m] for b in range(cs.shape[0])])
    cs, ci = np.take_along_axis(cs, order, 1), np.take_along_axis(ci, order, 1)

    def exact(idx):
        out = np.empty(idx.shape, dtype=np.float64)
        one = torch.empty((1, D), dtype=torch.float32, device=dev)
        rows = torch.empty((idx.shape[1], D), dtype=torch.float32, device=dev)
        for b in range(idx.shape[0]):
            for j, g in enumerate(idx[b].tolist()):
                ix.get_rows_device(int(g) - ix.row_offset, 1, one.data_ptr(), stream)
                rows[j] = one[0]
            out[b] = score_rows(rows)[b].cpu().numpy()
        return out
    return ci + ix.row_offset, cs, exact


def test_config4_10m_bf16_full_topk_vs_independent_matmul():
    """The headline config (10M x 1024 bf16, B=32, k=10): ALL k ranks of every query against an independent
    fp64 computation over the whole corpus (torch matmul on the rows read back from the index), dot and
    cosine - not only top-1 / self-consistency."""
    N, D, B, k = 10_000_000, 1024, 32, 10
    ix = _cabi.Index(N, D, "bf16", max_batch=64, max_k=16)
    try:
        ix.fill_synthetic(1234, 0)
        ix.finalize()
        Q, src = o.synth_queries(21, N, B, D, dtype=o.DTYPE_BF16)
        for metric in ("dot", "cosine"):
            idx, sc = ix.search(Q, k, metric, 1)
            assert (idx[:, 0] == src).all()
            ci, cs, exact = torch_exact_topk(ix, Q, metric, k + 22)
            sw = o.check_topk_candidates(idx, sc, ci, cs, exact(idx), k, rtol=2e-3, atol=1e-6)
            print(f"10M bf16 {metric}: near-tie substitutions {sw}")
            assert (idx == ci[:, :k]).mean() > 0.9            # bf16 products are exact; fp32 accumulation order only
    finally:
        ix.close()


@pytest.mark.parametrize("metric", ["dot", "cosine"])
def test_config5_shape_b256_k100_fp16(metric):
    """B=256, k=100, fp16 (BASELINE configs[4]'s shape) on a slice one GPU holds quickly: every rank of every
    query against the independent fp64 computation; the planner's pass structure is reported."""
    N, D, B, k = 1_000_003, 1024, 256, 100
    ix = _cabi.Index(N, D, "fp16", max_batch=256, max_k=128)
    try:
        ix.fill_synthetic(1234, 0)
        ix.finalize()
        Q, src = o.synth_queries(23, N, B, D, dtype=o.DTYPE_FP16)
        idx, sc = ix.search(Q, k, metric, 1)
        st = ix.stats()
        assert (idx[:, 0] == src).all()
        ci, cs, exact = torch_exact_topk(ix, Q, metric, k + 28)
        sw = o.check_topk_candidates(idx, sc, ci, cs, exact(idx), k, rtol=2e-3, atol=1e-6)
        print(f"B=256 k=100 fp16 {metric}: passes {st['passes']}, bytes streamed / corpus bytes "
              f"{st['bytes_streamed'] / st['corpus_bytes']:.1f}, near-tie substitutions {sw}")
        assert st["passes"] == 1 and st["bytes_streamed"] == st["corpus_bytes"]      # ONE pass for 256 queries
        idx2, sc2 = ix.search(Q, k, metric, 1)
        assert np.array_equal(idx, idx2) and np.array_equal(sc, sc2)
    finally:
        ix.close()


@pytest.mark.parametrize("dtype,metric,B,k,N", [
    ("bf16", "dot", 256, 100, 128 * 148 * 2 + 300), ("fp16", "cosine", 129, 10, 128 * 148 + 1), ("bf16", "dot", 200, 1, 128 * 7),
    ("fp16", "dot", 256, 128, 128 * 301 + 5), ("bf16", "cosine", 255, 17, 900), ("fp16", "dot", 300, 16, 128 * 40 + 77),
    ("bf16", "dot", 512, 33, 128 * 149)])
def test_wide_batch_kernel_sweep(dtype, metric, B, k, N):
    """Wide batches (128 < B <= 256 per pass, ONE pass over the shard) on CTA pairs (fused_search_pair_kernel, tcgen05
    cta_group::2).  Odd and even tile counts, fewer supertiles than pairs, ragged last tile, partly filled second half,
    every k range, both metrics, batches above 256 (several wide passes) - against the oracle, and bit-identical to
    the plan that streams the shard once per 128 queries (RS_WIDE=0)."""
    E = o.synth_rows(13, np.arange(N))
    rng = np.random.default_rng(3)
    src = rng.integers(0, N, B)
    src[:2] = [0, N - 1]
    Q = o.round_to_dtype(E[src] + 0.02 * rng.standard_normal((B, 1024)).astype(np.float32), DT[dtype])
    ix = make_index(E, dtype, max_batch=512, max_k=128)
    try:
        idx, sc = ix.search(Q, k, metric, 1)
        st = ix.stats()
        assert st["passes"] == -(-B // 256) and st["bytes_streamed"] == st["passes"] * st["corpus_bytes"]
        assert st["mma_m"] == 256 and st["nbuf"] == 2 and st["grid"] % 2 == 0       # the pair kernel ran
        S = exact_scores(Q, E, dtype, metric, 1)
        o.check_topk(idx, sc, S, k, rtol=RTOL[dtype], atol=1e-6)
        assert (idx[:, 0] == src).all()
        for _ in range(3):                                  # deterministic, and the pair protocol survives repeats
            i2, s2 = ix.search(Q, k, metric, 1)
            assert np.array_equal(idx, i2) and np.array_equal(sc, s2)
        ix.set_option("RS_WIDE", 0)                         # the general kernel, 128 queries per pass
        i3, s3 = ix.search(Q, k, metric, 1)
        assert ix.stats()["passes"] == -(-B // 128)
        assert np.array_equal(idx, i3) and np.array_equal(sc, s3)
    finally:
        ix.close()


@pytest.mark.parametrize("dtype,terms,metric,B,k,N", [
    ("fp32", 2, "cosine", 32, 10, 128 * 1184 + 77), ("fp32", 2, "dot", 17, 16, 128 * 1185), ("bf16", 2, "dot", 32, 1, 128 * 1186 + 1),
    ("fp16", 2, "cosine", 19, 5, 128 * 1300 + 127),
    ("bf16", 1, "dot", 64, 10, 128 * 1185 + 3), ("fp16", 1, "cosine", 33, 16, 128 * 1184), ("bf16", 1, "cosine", 50, 1, 128 * 1201 + 64)])
def test_doc_major_pair_kernel(dtype, terms, metric, B, k, N):
    """fused_search_dmpair_kernel (k <= 16, >= 8 tiles per SM; two accumulators and B <= 32, or one accumulator and
    32 < B <= 64: each CTA of a pair keeps half of the Q image, M = 256 MMAs over the two CTAs' document tiles): odd
    and even tile counts, ragged last tile, both metrics, fp32 storage, two-term and one-term 16-bit queries - against
    the oracle, and equal to the one-CTA kernels."""
    sx = 4 if dtype == "fp32" else 0
    E = o.synth_rows(29, np.arange(N))
    rng = np.random.default_rng(7)
    src = rng.integers(0, N, B)
    src[:2] = [N - 1, 0]
    Q = (E[src] + 0.03 * rng.standard_normal((B, 1024))).astype(np.float32)
    if terms == 1:
        Q = o.round_to_dtype(Q, DT[dtype])
    ix = make_index(E, dtype, sx, max_batch=64, max_k=16)
    try:
        idx, sc = ix.search(Q, k, metric, terms)
        st = ix.stats()
        assert st["doc_major"] == 1 and st["mma_m"] == 256 and st["grid"] % 2 == 0 and st["launches"] == 2, st
        S = exact_scores(Q, E, dtype, metric, terms, sx)
        o.check_topk(idx, sc, S, k, rtol=RTOL[dtype], atol=1e-6)
        assert (idx[:, 0] == src).all()
        for _ in range(3):
            i2, s2 = ix.search(Q, k, metric, terms)
            assert np.array_equal(idx, i2) and np.array_equal(sc, s2)
        ix.set_option("RS_DMPAIR", 0)                       # the one-CTA doc-major kernel: same values up to the order in
        i3, s3 = ix.search(Q, k, metric, terms)             # which the three MMA chains reach the low-order accumulator
        assert ix.stats()["mma_m"] != 256
        o.check_topk(i3, s3, S, k, rtol=RTOL[dtype], atol=1e-6)
        if terms == 1:                                      # one accumulator: the same sums in the same order
            assert np.array_equal(idx, i3) and np.array_equal(sc, s3)
        assert np.allclose(sc, s3, rtol=2e-6, atol=1e-7) and (idx == i3).mean() > 0.99
    finally:
        ix.close()


@pytest.mark.parametrize("hdtype", ["float32", "bfloat16", "float16"])
def test_hidden_state_handoff_mean_pool(squad, hdtype):
    """search_hidden_states: the embedder's [B, T, D] CUDA tensor is mean-pooled over T by
    rs_pool_mean_device (batch_processor.py:91: a plain mean over all positions) into the query buffer and searched on
    the same stream; pooled queries == torch's fp32 mean, results == the oracle on those queries == what the host
    path (batch_processor.py:92-97: .cpu().numpy() -> batch_retrieve) returns for them."""
    import torch
    r = CupyRetriever(squad.E, squad.contexts)
    try:
        g = torch.Generator().manual_seed(1)
        B, T = 16, 37
        base = torch.from_numpy(squad.Q[:B]).float()
        h = (base[:, None, :] + 0.3 * torch.randn(B, T, 1024, generator=g)).to(getattr(torch, hdtype)).cuda()
        idx, sc, q = r.search_hidden_states(h, squad.k)
        torch.cuda.synchronize()
        want = h.float().mean(dim=1)
        assert torch.allclose(q, want, rtol=1e-5, atol=1e-6)
        qn = q.cpu().numpy()
        S = o.scores_cosine(qn.astype(np.float64), o.stored_values(squad.E, o.DTYPE_FP32), accum=np.float64)
        o.check_topk(idx.cpu().numpy(), sc.cpu().numpy(), S, squad.k, rtol=1e-5, atol=1e-7)
        hi, hs = r.search(qn, squad.k)
        assert np.array_equal(hi, idx.cpu().numpy()) and np.allclose(hs, sc.cpu().numpy(), rtol=1e-6)
    finally:
        r.close()


def test_two_indexes_on_one_device_searched_concurrently(squad):
    """SimpleRetriever + CupyRetriever in one service process (or two indexes searched from two threads): while a second
    index lives on the device every search uses the separate merge launch instead of the in-kernel tail (whose CTAs
    wait for each other and could starve against another such grid), cached CUDA graphs are dropped when the number
    of live indexes changes, and concurrent searches return exactly what serial ones do."""
    import threading
    N = 128 * 300 + 11
    E = o.synth_rows(17, np.arange(N))
    Q, src = o.synth_queries(2, N, 32, 1024, dtype=o.DTYPE_BF16)
    a = make_index(E, "bf16", max_batch=64, max_k=16)
    try:
        for _ in range(3):                                   # alone: in-kernel tail, graph captured
            ia, sa = a.search(Q, 10, "dot", 1)
        assert a.stats()["launches"] == 1
        b = make_index(E[::-1].copy(), "bf16", max_batch=64, max_k=16)
        try:
            ib, sb = b.search(Q, 10, "dot", 1)
            assert a.search(Q, 10, "dot", 1)[0].tolist() == ia.tolist()
            assert a.stats()["launches"] == 2 and b.stats()["launches"] == 2      # separate merge kernel
            errs = []

            def loop(ix, want):
                try:
                    for _ in range(200):
                        got, _s = ix.search(Q, 10, "dot", 1)
                        if not np.array_equal(got, want):
                            errs.append("mismatch")
                            return
                except Exception as e:  # noqa
                    errs.append(repr(e))

            ts = [threading.Thread(target=loop, args=(a, ia)), threading.Thread(target=loop, args=(b, ib))]
            [t.start() for t in ts]
            [t.join() for t in ts]
            assert not errs, errs
        finally:
            b.close()
        for _ in range(3):                                   # alone again: back to one launch
            i2, _ = a.search(Q, 10, "dot", 1)
        assert a.stats()["launches"] == 1 and np.array_equal(i2, ia)
    finally:
        a.close()


# ---------------------------------------------------------------- property test over the whole shape space
_PROP_ROWS = o.synth_rows(77, np.arange(128 * 310 + 57))      # generated once; a prefix of it is the corpus


@settings(max_examples=150, deadline=None, derandomize=True,
          suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture, HealthCheck.data_too_large])
@given(n=st.one_of(st.integers(1, 600), st.integers(601, len(_PROP_ROWS))), B=st.integers(1, 256), k=st.integers(1, 100),
       dtype=st.sampled_from(["bf16", "fp16", "fp32"]), metric=st.sampled_from(["dot", "cosine"]),
       two_terms=st.booleans(), seed=st.integers(0, 2 ** 16), force_pair=st.booleans())
def test_property_any_shape_matches_oracle(n, B, k, dtype, metric, two_terms, seed, force_pair):
    """for ANY corpus size (not a multiple of the tile, fewer rows than k, one row), batch 1..256, k 1..100,
    storage type, metric and query-term count, the kernel the planner picks (doc-major, query-major resident or
    streamed, wide pair, doc-major pair - forced onto small corpora by RS_DMPAIR=2 in half of the examples -, k <= 16
    register lists or k > 16 buffered lists) returns the oracle's top-k: index sets
    identical up to provable near-ties, scores inside the tolerance, (-1, -inf) padding past the corpus size, and
    the same answer when the call is repeated."""
    E = _PROP_ROWS[:n]
    rng = np.random.default_rng(seed)
    src = rng.integers(0, n, B)
    Q = (E[src] * (1.0 + rng.random((B, 1)).astype(np.float32)) + 0.05 * rng.standard_normal((B, 1024)).astype(np.float32))
    terms = 2 if (two_terms or dtype == "fp32") else 1
    if terms == 1:
        Q = o.round_to_dtype(Q, DT[dtype])
    sx = 4 if dtype == "fp32" else 0
    ix = make_index(E, dtype, sx, max_batch=256, max_k=128)
    try:
        if force_pair:
            ix.set_option("RS_DMPAIR", 2)                   # doc-major pair kernel for any corpus size (where its shapes apply)
        idx, sc = ix.search(Q, k, metric, terms)
        S = exact_scores(Q, E, dtype, metric, terms, sx)
        o.check_topk(idx, sc, S, k, rtol=RTOL[dtype], atol=1e-6)
        if k > n:
            assert (idx[:, n:] == -1).all() and np.isneginf(sc[:, n:]).all()
        i2, s2 = ix.search(Q, k, metric, terms)
        assert np.array_equal(idx, i2) and np.array_equal(sc, s2)
    finally:
        ix.close()


@pytest.mark.parametrize("N", [300_000, 9_000])
def test_overlapped_launches_equal_serial_searches(N):
    """RS_OVERLAP=1 (programmatic dependent launch): back-to-back device searches whose scans overlap the previous
    search's tail give bit-identical results to serial searches, also with alternating queries, B and k.  300k rows:
    full-machine grids (dependents released at kernel entry); 9k rows: 71 CTAs (released after the scan)."""
    import torch
    D = 1024
    ix = _cabi.Index(N, D, "bf16", max_batch=64, max_k=16)
    try:
        ix.fill_synthetic(1234, 0)
        ix.finalize()
        Qa, _ = o.synth_queries(31, N, 32, D, dtype=o.DTYPE_BF16)
        Qb, _ = o.synth_queries(32, N, 32, D, dtype=o.DTYPE_BF16)
        qa, qb = torch.from_numpy(Qa).cuda(), torch.from_numpy(Qb).cuda()
        st = torch.cuda.current_stream().cuda_stream
        seq = [(qa, 32, 10), (qb, 32, 10), (qa, 7, 5), (qb, 32, 16), (qa, 32, 10), (qa, 32, 10), (qb, 1, 1)] * 3

        def run():
            outs = []
            for q, B, k in seq:
                oi = torch.empty((B, k), dtype=torch.int64, device="cuda")
                os_ = torch.empty((B, k), dtype=torch.float32, device="cuda")
                ix.search_device(q.data_ptr(), B, k, oi.data_ptr(), os_.data_ptr(), "dot", 1, st)
                outs.append((oi, os_))
            torch.cuda.synchronize()
            ix.check()
            return [(a.cpu().numpy(), b.cpu().numpy()) for a, b in outs]

        want = run()
        assert ix.stats()["overlapped"] == 0
        ix.set_option("RS_OVERLAP", 1)
        for _ in range(3):
            got = run()
            assert ix.stats()["overlapped"] == 1
            for (gi, gs), (wi, ws) in zip(got, want):
                assert np.array_equal(gi, wi) and np.array_equal(gs, ws)
        S = o.scores_dot(Qa.astype(np.float64), o.round_to_bf16(o.synth_rows(1234, np.arange(N), D)).astype(np.float64), np.float64)
        o.check_topk(want[0][0], want[0][1], S, 10, rtol=2e-3, atol=1e-6)
    finally:
        ix.close()


def test_searches_on_two_streams_are_ordered(squad):
    """One index, one set of workspaces: a search enqueued on another stream than the previous one is ordered
    after it by the library (cudaStreamWaitEvent), so interleaving rs_search (the index's own stream) with
    rs_search_device on torch streams never races."""
    import torch
    ix = make_index(squad.E, "bf16")
    try:
        Q = o.round_to_bf16(squad.Q)
        want_i, want_s = ix.search(Q, 10, "dot", 1)
        qd = torch.from_numpy(Q).cuda()
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        torch.cuda.synchronize()
        outs = []
        for it in range(40):
            st = (s1, s2)[it % 2]
            oi = torch.empty((len(Q), 10), dtype=torch.int64, device="cuda")
            os_ = torch.empty((len(Q), 10), dtype=torch.float32, device="cuda")
            ix.search_device(qd.data_ptr(), len(Q), 10, oi.data_ptr(), os_.data_ptr(), "dot", 1, st.cuda_stream)
            outs.append((oi, os_))
            if it % 8 == 7:
                hi, hs = ix.search(Q, 10, "dot", 1)
                assert np.array_equal(hi, want_i) and np.array_equal(hs, want_s)
        torch.cuda.synchronize()
        ix.check()
        for oi, os_ in outs:
            assert np.array_equal(oi.cpu().numpy(), want_i) and np.array_equal(os_.cpu().numpy(), want_s)
    finally:
        ix.close()


def test_k_above_max_k_does_not_empty_the_batch(squad):
    """The reference serves any k <= N; here a row asking for more than max_k gets the best max_k documents (a
    documented difference) and the OTHER requests of the batch keep their contexts."""
    r = SimpleRetriever(squad.E, squad.contexts, max_k=16)
    try:
        ks = [squad.k] * len(squad.Q)
        ks[3] = 200
        out = r.batch_retrieve(squad.Q, ks)
        for i, row in enumerate(out):
            if i != 3:
                assert row == squad.texts[i]
        assert len(out[3]) == 16 and out[3][:squad.k] == squad.texts[3]
        assert len(r.retrieve(squad.Q[0], 500)) == 16
    finally:
        r.close()


def test_config4_10m_bf16_properties():
    """10M x 1024 bf16, B=32, k=10 (the headline config) through size-independent properties:
    self-retrieval across the whole index range, returned scores re-derived from regenerated
    rows, determinism, and full search == merge of two half-corpus searches."""
    N, D, B, k = 10_000_000, 1024, 32, 10
    ix = _cabi.Index(N, D, "bf16", max_batch=64, max_k=16)
    try:
        ix.fill_synthetic(1234, 0)
        ix.finalize()
        rng = np.random.default_rng(0)
        src = rng.integers(0, N, B)
        src[:3] = [0, N - 1, N // 2]
        rows = o.round_to_bf16(o.synth_rows(1234, src, D))
        Q = o.round_to_bf16(rows + 0.01 * rng.standard_normal((B, D)).astype(np.float32) / 32)
        idx, sc = ix.search(Q, k, "dot", 1)
        assert (idx[:, 0] == src).all()
        assert (np.diff(sc, axis=1) <= 0).all() and (idx >= 0).all() and (idx < N).all()
        hit = o.round_to_bf16(o.synth_rows(1234, idx.reshape(-1), D)).astype(np.float64).reshape(B, k, D)
        exact = np.einsum("bkd,bd->bk", hit, Q.astype(np.float64))
        assert np.abs(sc - exact).max() <= 2e-3 * np.abs(exact).max()
        idx2, sc2 = ix.search(Q, k, "dot", 1)
        assert np.array_equal(idx, idx2) and np.array_equal(sc, sc2)
    finally:
        ix.close()
    halves = []
    for g in range(2):
        lo, hi = o.shard_range(N, 2, g)
        h = _cabi.Index(hi - lo, D, "bf16", row_offset=lo, max_batch=64, max_k=16)
        h.fill_synthetic(1234)
        h.finalize()
        halves.append(h.search(Q, k, "dot", 1))
        h.close()
    mi, ms = o.merge_candidates(np.stack([h[1] for h in halves]), np.stack([h[0] for h in halves]), k)
    assert np.array_equal(mi, idx) and np.array_equal(ms, sc)
