"""Small-corpus call latency through the host-buffer C ABI: the reference's real
corpora (100 / 1000 rows), B=32,k=10 and B=1,k=2 (service default), eager launches vs graph replay."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rag_serving_system_b200 import _cabi
from rag_serving_system_b200.retriever import CupyRetriever

G = os.path.join(ROOT, "tests", "golden")
for name in ("short_facts", "squad"):
    E = np.load(os.path.join(G, f"{name}_embeddings.npy"))
    Q = np.load(os.path.join(G, f"{name}_golden.npz"))["queries"]
    docs = json.load(open(os.path.join(G, f"{name}_contexts.json")))
    for dtype, metric, terms in (("bf16", "dot", 1), ("fp32", "cosine", 2)):
        for graph in ("0", "1"):
            os.environ["RS_GRAPH"] = graph
            ix = _cabi.Index(E.shape[0], E.shape[1], dtype, scale_exp=4 if dtype == "fp32" else 0, max_batch=64, max_k=16)
            ix.load_rows(E); ix.finalize()
            for B, k in ((32, 10), (1, 2)):
                q = Q[:B]
                for _ in range(20): ix.search(q, k, metric, terms)
                ts = []
                for _ in range(300):
                    t = time.perf_counter(); ix.search(q, k, metric, terms); ts.append(time.perf_counter() - t)
                ts = np.array(ts) * 1e6
                print(json.dumps({"corpus": name, "N": int(E.shape[0]), "dtype": dtype, "metric": metric, "B": B, "k": k, "graph": graph,
                                  "p50_us": round(float(np.median(ts)), 1), "p10_us": round(float(np.percentile(ts, 10)), 1),
                                  "p99_us": round(float(np.percentile(ts, 99)), 1)}), flush=True)
            ix.close()
    os.environ["RS_GRAPH"] = "1"
    r = CupyRetriever(E, docs)
    ks = [10] * 32
    for _ in range(20): r.batch_retrieve(Q, ks)
    ts = []
    for _ in range(300):
        t = time.perf_counter(); r.batch_retrieve(Q, ks); ts.append(time.perf_counter() - t)
    print(json.dumps({"corpus": name, "api": "CupyRetriever.batch_retrieve (strings out)", "B": 32, "k": 10,
                      "p50_us": round(float(np.median(ts)) * 1e6, 1)}), flush=True)
    r.close()
