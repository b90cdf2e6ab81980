"""Single-process multi-GPU (rs_group_*, behind B200Retriever(devices=[...])): per-call latency of search() and
batch_retrieve() with host buffers on the headline corpus, next to one GPU.

    python tools/group_probe.py [rows] [n_gpus]
"""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import retrieval_oracle as o
from rag_serving_system_b200 import B200Retriever

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
G = int(sys.argv[2]) if len(sys.argv) > 2 else torch.cuda.device_count()
B, k = 32, 10
Q, src = o.synth_queries(5, rows, B, 1024, dtype=o.DTYPE_BF16)
for devs in ([0], list(range(G))):
    r = B200Retriever.from_synthetic(rows, 1024, 1234, dtype="bf16", metric="dot", devices=devs, max_batch=64, max_k=16)
    for _ in range(5):
        idx, sc = r.search(Q, k)
    assert (idx[:, 0] == src).all()
    for name, fn in (("search", lambda: r.search(Q, k)), ("batch_retrieve", lambda: r.batch_retrieve(Q, [k] * B))):
        ts = []
        for _ in range(200):
            t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
        ts = np.array(ts) * 1e6
        print(json.dumps({"devices": len(devs), "rows": rows, "call": name, "p50_us": round(float(np.median(ts)), 1),
                          "p10_us": round(float(np.percentile(ts, 10)), 1), "p99_us": round(float(np.percentile(ts, 99)), 1),
                          "qps_p50": round(B / float(np.median(ts)) * 1e6, 1)}), flush=True)
    r.close()
