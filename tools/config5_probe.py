"""Config 5 (BASELINE.json configs[4]) per-GPU slice: 100M x 1024 fp16 over 8 GPUs = 12.5M rows per GPU,
B=256, k=100.  Prints kernel/step time, passes, q/s, achieved TFLOP/s and GB/s for a few plans.

    python tools/config5_probe.py [rows] [steps]
"""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from rag_serving_system_b200 import _cabi
from perf_sweep import run
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ix = _cabi.Index(rows, 1024, "fp16", max_batch=256, max_k=128)
ix.fill_synthetic(1234, 0); ix.finalize()
g = torch.Generator().manual_seed(0)
q = torch.randn(256, 1024, generator=g); q = (q / q.norm(dim=1, keepdim=True)).half().float().cuda()
out = []
cases = [(256, 100, {}), (256, 100, {"RS_BUFC": 32, "RS_LEAN": 0}), (128, 100, {}), (128, 100, {"RS_BUFC": 32, "RS_LEAN": 0}),
         (32, 100, {}), (256, 10, {}), (128, 10, {})]
for B, k, env in cases:
    run(ix, q, B, k, "dot", 1, steps, env, f"fp16 {rows} B{B} k{k}", out)
    r = out[-1]
    if "error" not in r:
        flops = 2.0 * B * 1024 * rows
        print(f"   -> {flops / (r['kernel_ms'] / 1e3) / 1e12:.0f} TFLOP/s useful, {r['passes']} pass(es), "
              f"{r['kernel_ms'] / r['passes']:.3f} ms/pass", flush=True)
