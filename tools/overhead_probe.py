"""Per-step overhead outside the fused kernel on one GPU, for a shard the size an 8-GPU run sees."""
import os, sys, json, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rag_serving_system_b200 import _cabi
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
ix = _cabi.Index(rows, 1024, "bf16", max_batch=64, max_k=16)
ix.fill_synthetic(1234, 0); ix.finalize()
g = torch.Generator().manual_seed(0)
q = torch.randn(32, 1024, generator=g); q = (q / q.norm(dim=1, keepdim=True)).bfloat16().float().cuda()
oi = torch.empty((32, 10), dtype=torch.int64, device="cuda"); os_ = torch.empty((32, 10), dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for env in ({}, {"RS_NO_EVENTS": "1"}):
    for k_ in ("RS_NO_EVENTS",): os.environ.pop(k_, None)
    os.environ.update(env)
    ix.reload_env_options()
    for _ in range(10): ix.search_device(q.data_ptr(), 32, 10, oi.data_ptr(), os_.data_ptr(), "dot", 1, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(steps): ix.search_device(q.data_ptr(), 32, 10, oi.data_ptr(), os_.data_ptr(), "dot", 1, st)
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    step = e0.elapsed_time(e1) / steps
    kt = float(ix.kernel_times_ms(min(steps, 256)).mean()) if not env else float("nan")
    print(json.dumps({"env": env, "rows": rows, "step_ms": round(step, 4), "kernel_ms": round(kt, 4), "overhead_us": round((step - kt) * 1e3, 1),
                      "cpu_issue_us_per_step": round((t1 - t0) / steps * 1e6, 1), "wall_us_per_step": round((t2 - t0) / steps * 1e6, 1)}), flush=True)
