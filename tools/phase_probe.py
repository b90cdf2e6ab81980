"""Phase breakdown of one fused launch from in-kernel %globaltimer stamps (RS_DEBUG_TS=1).

    python tools/phase_probe.py [rows] [B] [k] [dtype]
"""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["RS_DEBUG_TS"] = "1"
for a in [a for a in sys.argv[1:] if "=" in a]:      # NAME=VALUE arguments go to the environment
    os.environ[a.split("=")[0]] = a.split("=")[1]
    sys.argv.remove(a)
from rag_serving_system_b200 import _cabi
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dtype = sys.argv[4] if len(sys.argv) > 4 else "bf16"
ix = _cabi.Index(rows, 1024, dtype, max_batch=256, max_k=128, scale_exp=4 if dtype == "fp32" else 0)
ix.fill_synthetic(1234, 0); ix.finalize()
g = torch.Generator().manual_seed(0)
q = torch.randn(256, 1024, generator=g); q = (q / q.norm(dim=1, keepdim=True)).bfloat16().float().cuda()
oi = torch.empty((B, k), dtype=torch.int64, device="cuda"); os_ = torch.empty((B, k), dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
names = ["entry", "prep_done", "staged", "rounds_done", "arrived", "all_arrived", "merged"]
for it in range(6):
    ix.search_device(q.data_ptr(), B, k, oi.data_ptr(), os_.data_ptr(), "dot", 2 if dtype == "fp32" else 1, st)
    torch.cuda.synchronize()
    if it < 3:
        continue
    n = ix.stats()["grid"]
    ts = ix.debug_timestamps(n).astype(np.int64)
    t0 = ts[:, 0][ts[:, 0] > 0].min()
    rel = np.where(ts > 0, ts - t0, -1) / 1e3
    out = {"iter": it, "grid": n}
    for j, nm in enumerate(names):
        col = rel[:, j][rel[:, j] >= 0]
        if len(col):
            out[nm] = {"min": round(float(col.min()), 2), "med": round(float(np.median(col)), 2), "max": round(float(col.max()), 2)}
    cnt = ts[:, 7].astype(np.uint64)
    out["appends_per_cta"] = float((cnt & np.uint64(0xffffffff)).mean())
    out["flushes_per_cta"] = float((cnt >> np.uint64(32)).mean())
    print(json.dumps(out), flush=True)
