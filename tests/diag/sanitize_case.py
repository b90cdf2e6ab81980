"""Small searches for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck python tests/diag/sanitize_case.py
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import retrieval_oracle as o
from rag_serving_system_b200 import _cabi

rng = np.random.default_rng(0)
N, D = 128 * 150 + 37, 1024
E = o.synth_rows(3, np.arange(N), D)
for dtype, metric, B, k, terms in (("bf16", "dot", 32, 10, 1), ("fp32", "cosine", 7, 5, 2), ("fp16", "dot", 100, 40, 1),
                                   ("bf16", "cosine", 64, 16, 2), ("bf16", "dot", 3, 100, 1)):
    ix = _cabi.Index(N, D, dtype, max_batch=128, max_k=128, scale_exp=4 if dtype == "fp32" else 0)
    ix.load_rows(E); ix.finalize()
    src = rng.integers(0, N, B)
    Q = o.round_to_dtype(E[src] + 0.02 * rng.standard_normal((B, D)).astype(np.float32), o.DTYPE_NAMES[dtype])
    for _ in range(2):
        idx, sc = ix.search(Q, k, metric, terms)
    assert (idx[:, 0] == src).all(), (dtype, metric)
    print("ok", dtype, metric, B, k, ix.stats()["launches"], flush=True)
    ix.close()
