import json, os, sys
import torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
from rag_serving_system_b200 import _cabi
from perf_sweep import run
rows = 12_500_000
ix = _cabi.Index(rows, 1024, "fp16", max_batch=256, max_k=128)
ix.fill_synthetic(1234, 0); ix.finalize()
g = torch.Generator().manual_seed(0)
q = torch.randn(256, 1024, generator=g); q = (q / q.norm(dim=1, keepdim=True)).half().float().cuda()
out = []
for B, k, env in [(32, 10, {}), (32, 16, {}), (32, 17, {}), (32, 32, {}), (32, 64, {}), (32, 100, {}), (32, 100, {"RS_DEBUG_FLAGS": 2}), (32, 100, {"RS_CPS": 4}), (32,10,{"RS_CPS": 2}),
                  (128, 17, {}), (128, 100, {}), (128, 100, {"RS_DEBUG_FLAGS": 2}), (128, 10, {}), (128, 10, {"RS_DEBUG_FLAGS": 2})]:
    run(ix, q, B, k, "dot", 1, 10, env, f"B{B} k{k}", out)
    d = out[-1]
    print("  ", {kk: d.get(kk) for kk in ("label", "env", "kernel_ms", "stages", "sub", "nbuf", "smem", "sm_mhz", "error")}, flush=True)
