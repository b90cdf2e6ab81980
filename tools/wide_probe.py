"""Wide-batch kernel (B=256): where an epilogue warp's time goes (RS_DEBUG_TS cycle counters of warp 2, per CTA:
waiting for an accumulator / holding it / merging buffers afterwards) + appends and flushes per CTA.

    python tools/wide_probe.py [rows] [k] [NAME=VALUE ...]
"""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["RS_DEBUG_TS"] = "1"
for a in [a for a in sys.argv[1:] if "=" in a]:
    os.environ[a.split("=")[0]] = a.split("=")[1]
    sys.argv.remove(a)
from rag_serving_system_b200 import _cabi
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 100
B = 256
ix = _cabi.Index(rows, 1024, "fp16", max_batch=256, max_k=128)
ix.fill_synthetic(1234, 0); ix.finalize()
g = torch.Generator().manual_seed(0)
q = torch.randn(256, 1024, generator=g); q = (q / q.norm(dim=1, keepdim=True)).half().float().cuda()
oi = torch.empty((B, k), dtype=torch.int64, device="cuda"); os_ = torch.empty((B, k), dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for it in range(4):
    ix.search_device(q.data_ptr(), B, k, oi.data_ptr(), os_.data_ptr(), "dot", 1, st)
    torch.cuda.synchronize()
n = ix.stats()["grid"]
ts = ix.debug_timestamps(n).astype(np.int64)
kms = float(ix.kernel_times_ms(1)[0])
cnt = ts[:, 7].astype(np.uint64)
supertiles = -(-rows // 256) / n
out = {"rows": rows, "k": k, "kernel_ms": round(kms, 4), "supertiles_per_cta": round(supertiles, 1),
       "kcycles_per_supertile": {nm: round(float(np.median(ts[:, j])) / supertiles / 1e3, 2) for j, nm in ((1, "wait"), (2, "hold"), (3, "merge"))},
       "appends_per_cta": float((cnt & np.uint64(0xffffffff)).mean()), "flushes_per_cta": float((cnt >> np.uint64(32)).mean()),
       "stats": {kk: ix.stats()[kk] for kk in ("stages", "passes", "smem_bytes")}}
print(json.dumps(out))
