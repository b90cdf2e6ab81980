"""A/B of kernel knobs on one box: 10M x 1024 bf16, B=32, k=10."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from rag_serving_system_b200 import _cabi
from perf_sweep import run
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ix = _cabi.Index(rows, 1024, "bf16", max_batch=256, max_k=128)
ix.fill_synthetic(1234, 0); ix.finalize()
g = torch.Generator().manual_seed(0)
q = torch.randn(256, 1024, generator=g); q = (q / q.norm(dim=1, keepdim=True)).bfloat16().float().cuda()
out = []
for rep in range(2):
    for env in ({}, {"RS_CPS": 2}):
        run(ix, q, 32, 10, "dot", 1, steps, env, "B32", out)
for B in (1, 16):
    run(ix, q, B, 10, "dot", 1, steps, {}, f"B{B}", out)
for env in ({}, {"RS_QRESIDENT": 0}, {"RS_QRESIDENT": 0, "RS_SUB": 2}):
    run(ix, q, 64, 10, "dot", 1, steps, env, "B64", out)
for env in ({}, {"RS_SUB": 4}, {"RS_SUB": 1}, {"RS_CPS": 1, "RS_SUB": 4}):
    run(ix, q, 128, 10, "dot", 1, steps, env, "B128", out)
run(ix, q, 32, 10, "dot", 2, steps, {}, "B32 terms2", out)
run(ix, q, 32, 10, "dot", 2, steps, {"RS_QRESIDENT": 0}, "B32 terms2", out)
run(ix, q, 32, 10, "cosine", 1, steps, {}, "B32 cosine", out)
run(ix, q, 32, 100, "dot", 1, steps // 2, {}, "B32 k100", out)
run(ix, q, 128, 100, "dot", 1, steps // 2, {}, "B128 k100", out)
run(ix, q, 256, 10, "dot", 1, steps // 2, {}, "B256", out)
