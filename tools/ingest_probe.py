"""Ingest throughput: pageable numpy / memmapped .npy / pinned / device sources -> tiled HBM image."""
import json, os, sys, time, tempfile
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rag_serving_system_b200 import _cabi
N, D = 1_000_000, 1024
rng = np.random.default_rng(0)
E = rng.standard_normal((N, D), dtype=np.float32)
gb = E.nbytes / 1e9
for dtype in ("bf16", "fp32"):
    ix = _cabi.Index(N, D, dtype, scale_exp=0 if dtype != "fp32" else 1)
    for label, fn in (
        ("pageable fp32", lambda: ix.load_rows(E)),
        ("pinned fp32", None), ("device fp32", None)):
        if label == "pinned fp32":
            P = torch.from_numpy(E).pin_memory(); fn = lambda: ix.load_rows_ptr(P.data_ptr(), N, "fp32")
        if label == "device fp32":
            Dv = torch.from_numpy(E).cuda(); torch.cuda.synchronize(); fn = lambda: ix.load_rows_ptr(Dv.data_ptr(), N, "fp32")
        fn(); t = time.perf_counter(); fn(); el = time.perf_counter() - t
        print(json.dumps({"storage": dtype, "source": label, "GB": round(gb, 2), "s": round(el, 3), "GBps": round(gb / el, 1)}), flush=True)
    ix.close()
with tempfile.TemporaryDirectory() as d:
    np.save(os.path.join(d, "e.npy"), E)
    from rag_serving_system_b200 import B200Retriever
    t = time.perf_counter(); r = B200Retriever.from_npy(os.path.join(d, "e.npy"), dtype="bf16", metric="dot"); el = time.perf_counter() - t
    print(json.dumps({"storage": "bf16", "source": "np.load(mmap) via from_npy", "GB": round(gb, 2), "s": round(el, 3), "GBps": round(gb / el, 1)}))
    r.close()
