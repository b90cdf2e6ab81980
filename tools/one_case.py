"""One search configuration, a few steps (for ncu captures of a specific BASELINE config).

    python tools/one_case.py ROWS DTYPE B K [STEPS] [METRIC]
"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rag_serving_system_b200 import _cabi
rows, dtype, B, k = int(sys.argv[1]), sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 4
metric = sys.argv[6] if len(sys.argv) > 6 else "dot"
ix = _cabi.Index(rows, 1024, dtype, max_batch=256, max_k=128, scale_exp=4 if dtype == "fp32" else 0)
ix.fill_synthetic(1234, 0); ix.finalize()
g = torch.Generator().manual_seed(0)
q = torch.randn(256, 1024, generator=g); q = q / q.norm(dim=1, keepdim=True)
q = ({"bf16": q.bfloat16(), "fp16": q.half(), "fp32": q}[dtype]).float().cuda()
oi = torch.empty((B, k), dtype=torch.int64, device="cuda"); os_ = torch.empty((B, k), dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(steps):
    ix.search_device(q.data_ptr(), B, k, oi.data_ptr(), os_.data_ptr(), metric, 2 if dtype == "fp32" else 1, st)
torch.cuda.synchronize()
s = ix.stats()
kt = ix.kernel_times_ms(min(steps * s["passes"], 256))
print({"rows": rows, "dtype": dtype, "B": B, "k": k, "passes": s["passes"], "kernel_ms_per_pass": float(kt[-s["passes"]:].mean()), "stats": s})
