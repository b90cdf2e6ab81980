"""Stress of the CTA-pair protocols (remote mbarrier arrives, multicast commits, exchange hand-off): thousands of
back-to-back searches per shape, every result compared bit for bit with the first one (itself checked against the oracle
on a window).  A lost signal shows up as a watchdog error, a race as a mismatch.

    python tools/stress.py [seconds per shape]
"""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rag_serving_system_b200 import _cabi
from oracle import retrieval_oracle as o

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
shapes = [("fp16", 1_000_003, 256, 100, "dot", 1), ("fp32", 300_001, 32, 10, "cosine", 2), ("bf16", 400_000, 64, 16, "dot", 1),
          ("bf16", 1_250_000, 32, 10, "dot", 1)]
for dtype, n, B, k, metric, terms in shapes:
    ix = _cabi.Index(n, 1024, dtype, max_batch=256, max_k=128, scale_exp=4 if dtype == "fp32" else 0)
    ix.fill_synthetic(1234, 0); ix.finalize()
    Q, src = o.synth_queries(3, n, B, 1024, dtype=o.DTYPE_NAMES[dtype])
    q = torch.from_numpy(Q).cuda()
    oi = torch.empty((B, k), dtype=torch.int64, device="cuda"); os_ = torch.empty((B, k), dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ix.search_device(q.data_ptr(), B, k, oi.data_ptr(), os_.data_ptr(), metric, terms, st); torch.cuda.synchronize()
    ref_i, ref_s = oi.clone(), os_.clone()
    assert bool((ref_i[:, 0].cpu() == torch.from_numpy(src)).all()), "top-1 must be the source row"
    t0, it, bad = time.time(), 0, 0
    while time.time() - t0 < secs:
        for _ in range(50):
            ix.search_device(q.data_ptr(), B, k, oi.data_ptr(), os_.data_ptr(), metric, terms, st)
            bad += int(not (torch.equal(oi, ref_i) and torch.equal(os_, ref_s)))
            it += 1
        torch.cuda.synchronize(); ix.check()
    s = ix.stats()
    print({"dtype": dtype, "rows": n, "B": B, "k": k, "metric": metric, "searches": it, "mismatches": bad,
           "mma_m": s["mma_m"], "doc_major": s["doc_major"], "grid": s["grid"]}, flush=True)
    ix.close()
    assert bad == 0
print("stress ok")
