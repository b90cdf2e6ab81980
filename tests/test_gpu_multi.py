"""Multi-GPU parity (needs >= 2 B200s on the box; skipped otherwise)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


def test_sharded_search_under_torchrun():
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("multi-gpu ok") == 9
    assert res.stdout.count("overlapped launches ok") == 7        # every p2p case also runs with RS_OVERLAP=1


def test_single_process_multi_device_class(squad):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    from oracle import retrieval_oracle as o
    from rag_serving_system_b200 import B200Retriever
    r = B200Retriever(squad.E, squad.contexts, dtype="fp32", metric="dot", devices=list(range(min(n, 4))))
    idx, sc = r.search(squad.Q, squad.k)
    assert (idx == squad.idx_fixed).all()
    assert r.batch_retrieve(squad.Q, [squad.k] * len(squad.Q)) == squad.texts
    r.close()


def test_single_process_group_equals_single_gpu_and_oracle():
    """CupyRetriever-style object over several devices of ONE process (rs_group_*: one multi-device CUDA graph per
    shape, candidate exchange inside the kernels): identical to a single-GPU index of the same corpus and to
    the oracle, for the headline shape, a k > 16 shape, fp32 cosine, and ragged shards."""
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    from oracle import retrieval_oracle as o
    from rag_serving_system_b200 import B200Retriever, _cabi
    G = 8 if n >= 8 else (4 if n >= 4 else 2)
    N, D = 200_003, 1024
    for dtype, metric, B, k in (("bf16", "dot", 32, 10), ("fp16", "dot", 64, 100), ("fp32", "cosine", 9, 5)):
        multi = B200Retriever.from_synthetic(N, D, 1234, dtype=dtype, metric=metric, devices=list(range(G)))
        single = B200Retriever.from_synthetic(N, D, 1234, dtype=dtype, metric=metric, devices=[0])
        try:
            Q, src = o.synth_queries(41, N, B, D, dtype=o.DTYPE_NAMES[dtype])
            for it in range(4):                           # eager call, capture, two graph replays
                mi, ms = multi.search(Q, k)
            si, ss = single.search(Q, k)
            assert np.array_equal(mi, si) and np.array_equal(ms, ss)
            assert (mi[:, 0] == src).all()
            Q2 = Q[::-1].copy()                           # a replay must see THAT call's queries
            m2, _ = multi.search(Q2, k)
            s2, _ = single.search(Q2, k)
            assert np.array_equal(m2, s2)
            docs = multi.batch_retrieve(Q, [k] * B)
            assert docs == [[f"doc-{j}" for j in row] for row in mi.tolist()]
        finally:
            multi.close()
            single.close()
