"""CPU tests: the oracle against the reference-generated golden vectors (tests/golden, written by
oracle/make_golden.py from the UNMODIFIED reference SimpleRetriever) and its own invariants."""
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import retrieval_oracle as o

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---------------------------------------------------------------- pinned to the reference
def test_search_matches_reference_golden(golden):
    idx, val = o.search(golden.Q, golden.E, golden.k, "dot")
    assert (idx == golden.idx_fixed).all()
    # scores are what the reference computed them from (retriever.py:112)
    assert np.allclose(val, np.take_along_axis(golden.Q @ golden.E.T, idx, axis=1), rtol=1e-6)


def test_port_matches_reference_golden_variable_k(golden):
    r = o.SimpleRetrieverPort(golden.E, [str(i) for i in range(len(golden.E))])
    got = r.batch_retrieve(golden.Q, golden.ks_var)
    for row, kk, exp in zip(got, golden.ks_var, golden.idx_var):
        exp = [int(x) for x in exp if x >= 0]
        assert [int(s) for s in row] == exp
        assert len(row) == max(0, min(kk, len(golden.E)))


def test_port_texts_match_reference_golden(golden):
    r = o.SimpleRetrieverPort(golden.E, golden.contexts)
    assert r.batch_retrieve(golden.Q, [golden.k] * len(golden.Q)) == golden.texts


def test_self_retrieval_kat(golden):
    # query = corpus row i => top-1 = i, score 1 +- 1e-6, margin >= 0.048
    idx, val = o.search(golden.E, golden.E, 2, "dot")
    assert (idx[:, 0] == np.arange(len(golden.E))).all()
    assert (golden.self_top1 == np.arange(len(golden.E))).all()
    assert np.abs(val[:, 0] - 1.0).max() < 1e-5
    assert (val[:, 0] - val[:, 1]).min() > 0.04


def test_cosine_and_dot_rank_identically_on_unit_norm_corpora(golden):
    # SURVEY §0.3 [probed]: unit-normalised corpora => same ranking, different score values
    i_dot, _ = o.search(golden.Q, golden.E, golden.k, "dot")
    i_cos, v_cos = o.search(golden.Q, golden.E, golden.k, "cosine")
    assert (i_dot == i_cos).all()
    assert (np.abs(v_cos) <= 1.0).all()


def test_cosine_oracle_matches_reference_cupy_class_golden(golden):
    # goldens: the unmodified CupyRetriever (retriever.py:564-799) run with oracle/cupy_numpy_shim.py
    idx, val = o.search(golden.cos_Q, golden.E, golden.k, "cosine")
    assert (idx == golden.cos_idx_fixed).all()
    assert np.array_equal(val, golden.cos_scores_fixed)        # same fp32 operations in the same order
    assert not np.allclose(val, np.take_along_axis(golden.cos_Q @ golden.E.T, idx, axis=1), rtol=1e-2)
    docs = [str(i) for i in range(len(golden.E))]
    r = o.CupyRetrieverPort(golden.E, docs)
    got = r.batch_retrieve(golden.cos_Q, golden.ks_var)
    for row, exp in zip(got, golden.cos_idx_var):
        assert [int(s) for s in row] == [int(x) for x in exp if x >= 0]
    assert o.CupyRetrieverPort(golden.E, golden.contexts).batch_retrieve(
        golden.cos_Q, [golden.k] * len(golden.cos_Q)) == golden.cos_texts


def test_cupy_shim_runs_like_numpy():
    from oracle import cupy_numpy_shim as cp
    a = cp.asarray(np.array([[3.0, 1.0, 2.0]], dtype=np.float32))
    assert isinstance(a @ a.T, cp.ndarray) and isinstance(a[0], cp.ndarray)
    assert cp.argpartition(a[0], 1)[-2:].get().tolist() == [2, 0]       # full sort, like CuPy
    assert cp.argsort(-a[0]).get().tolist() == [0, 2, 1]
    assert type(cp.empty((2, 2), dtype=cp.int64).get()) is np.ndarray
    assert float(cp.linalg.norm(a, axis=1, keepdims=True)[0, 0]) == pytest.approx(14 ** 0.5)


def test_cupy_port_agrees_with_search(golden):
    docs = [str(i) for i in range(len(golden.E))]
    r = o.CupyRetrieverPort(golden.E, docs)
    got = r.batch_retrieve(golden.Q, [golden.k] * len(golden.Q))
    idx, _ = o.search(golden.Q, golden.E, golden.k, "cosine")
    assert [[int(s) for s in row] for row in got] == idx.tolist()
    assert r.retrieve(golden.Q[0], 3) == [str(i) for i in idx[0, :3]]
    assert r.retrieve(golden.Q[0].reshape(1, -1), 3) == [str(i) for i in idx[0, :3]]


def test_cupy_port_return_conventions(short_facts):
    g = short_facts
    r = o.CupyRetrieverPort(g.E, g.contexts)
    B = 4
    Q = g.Q[:B]
    assert r.batch_retrieve(Q[0], [1]) == [[]]                        # ndim != 2        (:708-710)
    assert r.batch_retrieve(Q[:0], []) == []                          # empty batch      (:711-713)
    assert r.batch_retrieve(Q, [1, 2]) == [[]] * B                    # len mismatch     (:714-716)
    assert r.batch_retrieve(Q, (1, 1, 1, 1)) == [[]] * B              # ks not a list    (:714)
    assert r.batch_retrieve(Q[:, :8], [1] * B) == [[]] * B            # D mismatch       (:717-719)
    assert r.batch_retrieve(Q, [1, 1, np.int64(1), 1]) == [[]] * B    # non-int k        (:720-722)
    assert r.batch_retrieve(Q, [0, -1, 0, 0]) == [[]] * B             # no positive k    (:726-728)
    out = r.batch_retrieve(Q, [0, 3, 1000, 1])                        # clamp to N, k<=0 -> []
    assert [len(x) for x in out] == [0, 3, len(g.E), 1]
    assert r.retrieve("nope", 1) == [] and r.retrieve(Q, 1) == [] and r.retrieve(Q[0], 0) == []
    assert len(r.retrieve(Q[0], 10 ** 6)) == len(g.E)


# ---------------------------------------------------------------- tie contract / merge / shards
def test_ties_go_to_lower_index():
    D = np.zeros((6, 4), np.float32)
    D[[0, 3, 5]] = [1, 0, 0, 0]       # three identical best rows
    D[[1, 2]] = [0.5, 0, 0, 0]
    Q = np.array([[1, 0, 0, 0]], np.float32)
    idx, val = o.search(Q, D, 4, "dot")
    assert idx.tolist() == [[0, 3, 5, 1]]
    idx, val = o.search(Q, D, 2, "dot")
    assert idx.tolist() == [[0, 3]]
    idx, _ = o.search(Q, np.ones((5, 4), np.float32), 3, "cosine")     # all-equal rows
    assert idx.tolist() == [[0, 1, 2]]


def test_topk_pads_when_k_exceeds_n():
    idx, val = o.search(np.ones((2, 4), np.float32), np.eye(4, dtype=np.float32)[:3], 5, "dot")
    assert (idx[:, 3:] == -1).all() and np.isinf(val[:, 3:]).all() and (idx[:, :3] >= 0).all()


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 300), st.integers(1, 7), st.integers(1, 12), st.integers(0, 2 ** 31))
def test_sharded_merge_equals_global(n, g, k, seed):
    rng = np.random.default_rng(seed)
    D = rng.integers(-3, 4, size=(n, 8)).astype(np.float32)           # many exact ties
    Q = rng.integers(-3, 4, size=(5, 8)).astype(np.float32)
    gi, gs = o.search(Q, D, k, "dot")
    cs, ci = [], []
    for r in range(g):
        lo, hi = o.shard_range(n, g, r)
        i, s = o.search(Q, D[lo:hi], k, "dot") if hi > lo else (np.full((5, k), -1), np.full((5, k), -np.inf))
        ci.append(np.where(i >= 0, i + lo, -1))
        cs.append(s)
    mi, ms = o.merge_candidates(np.stack(cs), np.stack(ci), k)
    assert (mi == gi).all() and np.array_equal(ms, gs)


def test_shard_range_partitions():
    for n in (1, 7, 128, 1000, 10_000_000):
        for g in (1, 2, 3, 4, 8):
            spans = [o.shard_range(n, g, r) for r in range(g)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_check_topk_rejects_wrong_answers():
    rng = np.random.default_rng(0)
    S = rng.standard_normal((3, 50))
    idx, val = o.topk_rows(S.astype(np.float32), 5)
    o.check_topk(idx, val, S, 5, rtol=1e-6, atol=1e-7)
    bad = idx.copy()
    bad[0, 4] = int(np.argmin(S[0]))
    with pytest.raises(AssertionError):
        o.check_topk(bad, val, S, 5, rtol=1e-6, atol=1e-7)
    with pytest.raises(AssertionError):
        o.check_topk(idx, val + 0.1, S, 5, rtol=1e-6, atol=1e-7)


# ---------------------------------------------------------------- dtype helpers
def test_bf16_rounding_matches_torch():
    import torch
    x = np.random.default_rng(1).standard_normal(10000).astype(np.float32) * 3
    x[:4] = [0.0, -0.0, 1.0000001, 65504.0]
    assert np.array_equal(o.round_to_bf16(x), torch.from_numpy(x).bfloat16().float().numpy())
    assert np.array_equal(o.bf16_bits_to_f32(o.bf16_bits(x)), o.round_to_bf16(x))


def test_fp16_split_precision():
    x = (np.random.default_rng(2).standard_normal(100000) / 32).astype(np.float32)
    for e in (0, 4):
        hi, lo = o.split_fp16(x, e)
        rep = o.unsplit_fp16(hi, lo, e)
        rel = np.abs(rep - x.astype(np.float64)) / np.maximum(np.abs(x), 1e-4)
        assert rel.max() < 2.0 ** -21, rel.max()
    hi, lo = o.split_bf16(x)
    assert (np.abs((hi + lo).astype(np.float64) - x) <= np.abs(x) * 2.0 ** -16).all()


# ---------------------------------------------------------------- synthetic corpus / layout
def test_synth_rows_deterministic_and_unit_norm():
    a = o.synth_rows(1234, np.arange(100, 140))
    b = o.synth_rows(1234, np.arange(0, 200))[100:140]
    assert np.array_equal(a, b)                                        # any shard reproduces the same rows
    assert np.abs(np.linalg.norm(a.astype(np.float64), axis=1) - 1).max() < 1e-6
    assert not np.array_equal(a, o.synth_rows(1235, np.arange(100, 140)))
    assert abs(float(a.mean())) < 1e-2
    # known-answer vector (guards the generator both sides of the ABI share)
    assert np.allclose(o.synth_rows(1234, np.array([0]))[0, :4],
                       [-0.04290489, -0.0350121, 0.05111211, 0.02555394], atol=1e-8)


def test_tile_layout_is_a_bijection_with_swizzle():
    off = o.tile_layout_offsets(300, 128, planes=2)
    flat = np.sort(off.reshape(-1))
    assert len(np.unique(flat)) == flat.size and (flat % 2 == 0).all()
    # rows >= 300 of the last tile are the only holes
    n_tiles, tile_bytes = 3, 2 * 2 * o.BLOCK_BYTES
    assert flat.max() < n_tiles * tile_bytes
    # spot checks of the Swizzle<3,4,3> rule: row r, 16-byte chunk c lands at chunk c ^ (r % 8)
    o1 = o.tile_layout_offsets(16, 64)[0]
    assert o1[0, 0] == 0 and o1[0, 8] == 16 and o1[1, 0] == 128 + 16 and o1[1, 8] == 128
    assert o1[9, 0] == 1024 + 128 + 16 and o1[7, 56] == 7 * 128 + 0
    img = o.tile_corpus(np.arange(2 * 130 * 64, dtype=np.uint16).reshape(2, 130, 64))
    assert img.size == 2 * 2 * o.BLOCK_BYTES // 2


def test_stored_values_roundtrip():
    x = o.synth_rows(9, np.arange(10))
    assert np.array_equal(o.stored_values(x, o.DTYPE_BF16), o.round_to_bf16(x).astype(np.float64))
    assert np.abs(o.stored_values(x, o.DTYPE_FP32, 4) - x).max() < 1e-8


def test_native_synth_rows_equal_numpy():
    """oracle/synth_native.c (the C restatement bench.py's CPU legs use for >= 1M-row samples) is bit-identical to
    the NumPy generator, for any row range and dimension."""
    from oracle import build_oracle
    if build_oracle.build_native() is None:
        pytest.skip("gcc not available")
    for seed, row0, n, D in ((1234, 0, 300, 1024), (7, 999_999_000, 129, 64), (1234, 12_345_678, 50, 2048)):
        a = build_oracle.synth_rows_native(seed, row0, n, D)
        assert a.dtype == np.float32 and np.array_equal(a, o.synth_rows(seed, np.arange(row0, row0 + n), D))


def test_reference_copy_recipe_and_port_agree():
    """oracle/_ref holds an UNMODIFIED copy of the reference module (made by oracle/build_oracle.py when
    /root/reference exists); its SimpleRetriever and the oracle port return the same documents."""
    import hashlib
    from oracle import build_oracle
    mod = build_oracle.load_reference_module()
    if mod is None:
        pytest.skip("no /root/reference and no oracle/_ref copy on this box")
    if os.path.exists(build_oracle.REF_SRC):
        assert (hashlib.sha256(open(build_oracle.REF_SRC, "rb").read()).hexdigest()
                == hashlib.sha256(open(build_oracle.REF_COPY, "rb").read()).hexdigest())
    g = np.load(os.path.join(GOLD, "squad_golden.npz"))
    E = np.load(os.path.join(GOLD, "squad_embeddings.npy"))
    docs = [str(i) for i in range(len(E))]
    ks = [int(g["k"])] * len(g["queries"])
    assert mod.SimpleRetriever(E, docs).batch_retrieve(g["queries"], ks) == o.SimpleRetrieverPort(E, docs).batch_retrieve(g["queries"], ks)
