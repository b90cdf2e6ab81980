"""Where does the power go?  Same 10M x 1024 bf16 scan with parts of the kernel disabled
(RS_DEBUG_FLAGS: 1 = no MMAs, 2 = no epilogue; results are garbage, only timing is read)."""
import os, sys, json, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rag_serving_system_b200 import _cabi
sys.path.insert(0, os.path.join(ROOT, "tools"))
from perf_sweep import run

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 150
ix = _cabi.Index(rows, 1024, "bf16", max_batch=256, max_k=128)
ix.fill_synthetic(1234, 0); ix.finalize()
g = torch.Generator().manual_seed(0)
q = torch.randn(256, 1024, generator=g); q = (q / q.norm(dim=1, keepdim=True)).bfloat16().float().cuda()
out = []
for rep in range(2):
    for flags in (0, 1, 3, 2, 0):
        os.environ["RS_DEBUG_FLAGS"] = str(flags)
        run(ix, q, 32, 10, "dot", 1, steps, {"RS_MMA_M": 64}, f"flags={flags}", out)
        if flags == 0:
            run(ix, q, 32, 10, "dot", 1, steps, {"RS_MMA_M": 128}, f"flags={flags} M128", out)
os.environ["RS_DEBUG_FLAGS"] = "0"
# a plain device-to-device copy and a read-only reduction of the same bytes for reference
x = torch.empty(rows * 1024, dtype=torch.bfloat16, device="cuda"); y = torch.empty_like(x)
for name, fn, nbytes in (("torch copy (r+w)", lambda: y.copy_(x), 2 * x.numel() * 2), ("torch sum (r)", lambda: x.view(torch.int32).sum(), x.numel() * 2)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"label": name, "ms": round(ms, 4), "GBps": round(nbytes / ms / 1e6, 1)}), flush=True)
