"""CPU tests of the host side above the C ABI: the retriever class's reference-compatible
conventions (with the device index replaced by an oracle-backed fake - a test double, not a
product path) and the multi-rank candidate exchange over gloo with world_size 2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import retrieval_oracle as o
from rag_serving_system_b200 import _cabi, retriever, sharded


class FakeIndex:
    """Test double for _cabi.Index: same surface, answers come from the oracle."""

    def __init__(self, n_rows, dim, dtype="bf16", device=0, row_offset=0, max_batch=256, max_k=128, scale_exp=0):
        self.n_rows, self.dim, self.dtype, self.device, self.row_offset = n_rows, dim, dtype, device, row_offset
        self.rows = np.zeros((n_rows, dim), np.float32)
        self.finalized = False

    def load_rows(self, rows, row0=0, src_dtype=None):
        self.rows[row0:row0 + len(rows)] = rows

    def finalize(self):
        self.finalized = True

    def search(self, q, k, metric="dot", query_terms=1):
        assert self.finalized
        i, s = o.search(q, self.rows, k, metric)
        return np.where(i >= 0, i + self.row_offset, -1), s

    def close(self):
        pass


class FakeGroup:
    """Test double for _cabi.Group: per-shard oracle searches merged with the oracle's candidate merge."""

    def __init__(self, shards):
        self.shards = list(shards)

    def search(self, q, k, metric="dot", query_terms=1):
        parts = [ix.search(q, k, metric, query_terms) for ix in self.shards]
        return o.merge_candidates(np.stack([p[1] for p in parts]), np.stack([p[0] for p in parts]), k)

    def close(self):
        pass


@pytest.fixture()
def fake_backend(monkeypatch):
    monkeypatch.setattr(_cabi, "Index", FakeIndex)
    monkeypatch.setattr(_cabi, "Group", FakeGroup)


def test_class_surface_matches_reference(fake_backend, short_facts):
    g = short_facts
    r = retriever.CupyRetriever(g.E, g.contexts)
    assert (r.N_A, r.D) == g.E.shape and r.documents is g.contexts       # retriever.py:589-591
    ref = o.CupyRetrieverPort(g.E, g.contexts)
    ks = [g.k] * len(g.Q)
    assert r.batch_retrieve(g.Q, ks) == ref.batch_retrieve(g.Q, ks)
    assert r.batch_retrieve(g.Q, g.ks_var) == ref.batch_retrieve(g.Q, g.ks_var)
    assert r.retrieve(g.Q[3], 4) == ref.retrieve(g.Q[3], 4)
    assert r.retrieve(g.Q[3].reshape(1, -1)) == ref.retrieve(g.Q[3])     # default k=2, (1,D) accepted
    s = retriever.SimpleRetriever(g.E, g.contexts)
    assert s.batch_retrieve(g.Q, ks) == g.texts                          # == unmodified reference output


def test_error_conventions_never_raise(fake_backend, short_facts):
    g = short_facts
    r = retriever.CupyRetriever(g.E, g.contexts)
    ref = o.CupyRetrieverPort(g.E, g.contexts)
    Q = g.Q[:4]
    cases = [
        (Q[0], [1]), (Q[:0], []), (Q, [1, 2]), (Q, (1, 1, 1, 1)), (Q[:, :8], [1] * 4),
        (Q, [1, 1, np.int64(1), 1]), (Q, [0, -1, 0, 0]), (Q, [0, 3, 1000, 1]), (Q.astype(np.float64), [2] * 4),
    ]
    for q, ks in cases:
        assert r.batch_retrieve(q, ks) == ref.batch_retrieve(q, ks), (q.shape, ks)
    assert r.batch_retrieve("nope", [1, 2]) == [[], []]
    for q, k in (("nope", 1), (Q, 1), (Q[0], 0), (Q[0], -3), (Q[0, :5], 1), (Q[0], 10 ** 6)):
        assert r.retrieve(q, k) == ref.retrieve(q, k)


def test_retrieve_with_scores_triple_and_multi_shard(fake_backend, squad):
    g = squad
    r = retriever.B200Retriever(g.E, g.contexts, metric="dot", dtype="fp32")
    idx, sc, ctx = r.retrieve_with_scores(g.Q, g.k)
    assert (idx == g.idx_fixed).all() and ctx == g.texts and sc.shape == idx.shape
    with pytest.raises(ValueError):
        r.search(g.Q, 0)
    with pytest.raises(ValueError):
        r.search(g.Q[:, :3], 1)


def test_multi_device_class_and_k_above_max_k(fake_backend, squad):
    """devices=[...] splits the corpus into contiguous row shards behind one group (same answers as one shard);
    a request with k > max_k gets the best max_k documents and does not empty the other rows of its batch
    (the reference serves any k <= N, retriever.py:729-730; documented difference)."""
    g = squad
    one = retriever.SimpleRetriever(g.E, g.contexts, max_k=16)
    multi = retriever.SimpleRetriever(g.E, g.contexts, max_k=16, devices=[0, 1, 2])
    assert [ix.row_offset for ix in multi._shards] == [0, 334, 668] and multi._group is not None
    ks = [g.k] * len(g.Q)
    assert multi.batch_retrieve(g.Q, ks) == one.batch_retrieve(g.Q, ks) == g.texts
    ks[5] = 99
    out = one.batch_retrieve(g.Q, ks)
    assert [row for i, row in enumerate(out) if i != 5] == [row for i, row in enumerate(g.texts) if i != 5]
    assert len(out[5]) == 16 and out[5][:g.k] == g.texts[5]
    assert len(one.retrieve(g.Q[0], 40)) == 16 and one.retrieve(g.Q[0], 3) == g.texts[0][:3]


def test_gather_and_source_dtype_helpers():
    docs = [f"d{i}" for i in range(50)]
    assert retriever._gather(docs, np.array([4, 0, 49])) == ["d4", "d0", "d49"]
    assert retriever._gather(docs, np.array([7])) == ["d7"] and retriever._gather(docs, np.array([], np.int64)) == []
    lazy = retriever._LazyDocs(10)
    assert retriever._gather(lazy, np.array([2, 9])) == ["doc-2", "doc-9"] and len(lazy) == 10
    with pytest.raises(IndexError):
        lazy[10]
    assert retriever._source_dtype(np.dtype(np.float32), None) == "fp32"
    assert retriever._source_dtype(np.dtype(np.float16), None) == "fp16"
    assert retriever._source_dtype(np.dtype("V2"), None) == "bf16"
    assert retriever._source_dtype(np.dtype(np.float64), None) is None
    assert retriever._source_dtype(np.dtype(np.uint16), "bf16") == "bf16"
    with pytest.raises(ValueError):
        retriever._source_dtype(np.dtype(np.float32), "bf16")
    x = (np.arange(12, dtype=np.float32).reshape(3, 4) - 5) * 0.25
    assert retriever._max_abs_chunked(x, "fp32") == 1.5
    assert retriever._max_abs_chunked(o.bf16_bits(x), "bf16") == 1.5


def test_launch_planner_is_consistent():
    """rs_debug_plan (host only): every plan fits the TMEM / shared-memory budget the kernel assumes."""
    for dim in (64, 128, 1024, 2048, 4096):
        for dtype, terms in (("bf16", 1), ("bf16", 2), ("fp16", 1), ("fp32", 2)):
            for B in (1, 8, 31, 32, 33, 64, 65, 100, 128, 129, 256):
                for k in (1, 10, 16, 17, 100, 128):
                    try:
                        pl = _cabi.debug_plan(dim, dtype, B, k, terms)
                    except _cabi.RsError:
                        assert dim >= 2048 or k > 16, (dim, dtype, B, k)      # only big shapes may be refused
                        continue
                    nacc = 2 if (terms == 2 or dtype == "fp32") else 1
                    if B > 128 and nacc == 1:            # wide-batch (CTA-pair) kernel: both 128-row halves share ONE pass
                        assert (pl["passes"], pl["bpad"], pl["sub"], pl["nbuf"], pl["mma_m"], pl["cps"], pl["q_resident"]) == \
                            (1, 256, 2, 2, 256, 1, 0), (dim, dtype, B, k, pl)
                        continue
                    assert pl["passes"] * pl["per_pass"] >= B and pl["bpad"] % 8 == 0 and pl["bpad"] <= 128
                    assert pl["bpad"] >= pl["per_pass"] and pl["stages"] >= 2
                    assert pl["sub"] * pl["nbuf"] * nacc in (1, 2, 4)
                    assert pl["mma_m"] == (64 if pl["bpad"] <= 64 else 128)
                    assert (dim // 64) % pl["cps"] == 0
                    if pl["q_resident"]:
                        assert pl["sub"] == 1 and pl["nbuf"] == 4 // nacc
    hl = _cabi.debug_plan(1024, "bf16", 32, 10, 1)                            # the headline config
    assert hl == {"passes": 1, "per_pass": 32, "bpad": 32, "stages": 2, "q_resident": 1, "sub": 1, "nbuf": 4,
                  "mma_m": 64, "cps": 4}


def test_auto_query_terms():
    q = np.random.default_rng(0).standard_normal((4, 64)).astype(np.float32)
    assert retriever._auto_query_terms(q, "bf16") == 2
    assert retriever._auto_query_terms(o.round_to_bf16(q), "bf16") == 1
    assert retriever._auto_query_terms(q, "fp16") == 2
    assert retriever._auto_query_terms(o.round_to_fp16(q), "fp16") == 1
    assert retriever._auto_query_terms(o.round_to_fp16(q), "fp32") == 2


def test_fp32_scale_exp_keeps_split_in_fp16_range():
    for m in (1e-3, 0.07, 1.0, 37.5, 6e4):
        e = retriever._fp32_scale_exp(m)
        assert 8.0 <= m * 2.0 ** e < 16.0
    assert retriever._fp32_scale_exp(0.0) == 0


# ---------------------------------------------------------------- N>1 exchange step over gloo
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, n, B, k, seed, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(seed)
        D = rng.integers(-3, 4, size=(n, 16)).astype(np.float32)
        Q = rng.integers(-3, 4, size=(B, 16)).astype(np.float32)
        lo, hi = sharded.shard_range(n, world, rank)
        i, s = o.search(Q, D[lo:hi], k, "dot")                 # stands in for the local fused kernel
        i = np.where(i >= 0, i + lo, -1)
        nb = sharded.packed_bytes(B, k)
        local = torch.zeros(nb, dtype=torch.uint8)
        li, ls = sharded.pack_views(local, B, k)
        li.copy_(torch.from_numpy(i))
        ls.copy_(torch.from_numpy(s))
        gathered = torch.empty((world, nb), dtype=torch.uint8)
        sharded.gather_candidates(local, gathered)
        ci = np.stack([sharded.pack_views(gathered[g], B, k)[0].numpy() for g in range(world)])
        cs = np.stack([sharded.pack_views(gathered[g], B, k)[1].numpy() for g in range(world)])
        mi, ms = o.merge_candidates(cs, ci, k)                 # stands in for rs_merge_device
        gi, gs = o.search(Q, D, k, "dot")
        ret[rank] = bool((mi == gi).all() and np.array_equal(ms, gs))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,B,k", [(1000, 8, 10), (37, 3, 7), (5, 1, 1)])
def test_candidate_exchange_world2_gloo(n, B, k):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_rank_main, args=(world, _free_port(), n, B, k, 7, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_packed_buffer_alignment():
    for B, k in ((1, 1), (3, 7), (32, 10), (256, 100)):
        nb = sharded.packed_bytes(B, k)
        assert nb % 16 == 0 and nb >= B * k * 12
        buf = torch.zeros(nb, dtype=torch.uint8)
        i, s = sharded.pack_views(buf, B, k)
        assert i.shape == (B, k) and s.shape == (B, k) and i.dtype == torch.int64 and s.dtype == torch.float32


def test_rows_to_docs_matches_reference_slicing():
    """The batch helper behind batch_retrieve (retriever.py:776-790): same lists as the per-row
    ``row[row >= 0][:k_i]`` + lookup it replaces, for padded rows, k_i <= 0, k_i > row length, tuple docs."""
    from rag_serving_system_b200.retriever import _rows_to_docs
    rng = np.random.default_rng(0)
    docs = [f"d{i}" for i in range(50)]
    for trial in range(200):
        B, k = int(rng.integers(1, 9)), int(rng.integers(1, 12))
        idx = rng.integers(0, 50, size=(B, k)).astype(np.int64)
        if trial % 2:                                   # -1 padding at the end of some rows (N < k)
            for r in range(B):
                n_pad = int(rng.integers(0, k + 1))
                if n_pad:
                    idx[r, k - n_pad:] = -1
        ks = [int(x) for x in rng.integers(-2, k + 3, size=B)]
        want = []
        for i in range(B):
            if ks[i] <= 0:
                want.append([])
            else:
                row = idx[i]
                want.append([docs[int(j)] for j in row[row >= 0][:ks[i]]])
        assert _rows_to_docs(docs, idx, ks) == want
        assert _rows_to_docs(tuple(docs), idx, ks) == want
    got = _rows_to_docs(docs, np.array([[3, 4]], dtype=np.int64), [2])
    got[0].append("x")                                  # fresh lists: mutating the result never touches documents
    assert docs[3] == "d3" and len(docs) == 50


def test_document_arena_join_matches_reference_joins(squad):
    """rs_docs_* (host code of the C ABI, no GPU): per query the first k_i valid hits joined with the separator of
    batch_processor.py:100 == "\\n---\\n".join(documents[i] ...) of the reference's slicing (retriever.py:776-790);
    padding (-1), k_i <= 0, k_i > stride, non-ASCII and empty documents."""
    docs = list(squad.contexts[:200]) + ["", "éü中文 \U0001f600", "tail"]
    arena = _cabi.DocArena(docs)
    rng = np.random.default_rng(1)
    idx = rng.integers(0, len(docs), (16, 7))
    idx[2, 4:] = -1
    idx[3, :] = -1
    idx[4, :3] = [200, 201, 202]
    ks = [7, 2, 7, 5, 3, 0, -4, 100, 1, 7, 7, 6, 5, 4, 3, 2]
    want = ["\n---\n".join(docs[j] for j in [x for x in row if x >= 0][:max(0, k)]) for row, k in zip(idx.tolist(), ks)]
    assert arena.join(idx, ks) == want
    assert arena.join(idx, ks, sep=" | ") == [w.replace("\n---\n", " | ") for w in want]
    assert arena.join(np.zeros((0, 3), np.int64), []) == []
    with pytest.raises(_cabi.RsError):
        arena.join(np.array([[len(docs)]]), [1])            # an index outside the corpus is an error, not a crash
    arena.close()


def test_batch_retrieve_contexts_equals_joined_batch_retrieve(fake_backend, squad):
    """batch_retrieve_contexts == [sep.join(d) for d in batch_retrieve(...)] (batch_processor.py:95-100), including
    the reference's error conventions, and assemble_prompts == the template of batch_processor.py:103."""
    r = retriever.CupyRetriever(squad.E, squad.contexts)
    ks = list(squad.ks_var)
    want = ["\n---\n".join(d) for d in r.batch_retrieve(squad.cos_Q, ks)]
    assert r.batch_retrieve_contexts(squad.cos_Q, ks) == want
    fixed = r.batch_retrieve_contexts(squad.cos_Q, [squad.k] * len(squad.cos_Q))
    assert fixed == ["\n---\n".join(t) for t in squad.cos_texts]      # strings of the unmodified reference class, joined
    assert r.batch_retrieve_contexts(squad.cos_Q, [0] * len(ks)) == [""] * len(ks)
    assert r.batch_retrieve_contexts(squad.cos_Q[:, :5], ks) == [""] * len(ks)          # wrong dimension: [] rows
    assert r.batch_retrieve_contexts("nope", ks) == [""] * len(ks)
    qs = [f"q{i}" for i in range(len(ks))]
    prompts = r.assemble_prompts(qs, want)
    assert prompts[3] == f"Context:\n{want[3]}\n\nQuestion: q3\n\nThe Answer to this question is: "
    r.close()
