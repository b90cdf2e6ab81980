"""Kernel-variant sweep on one B200 (run under gpurun).  Prints one line per variant with the fused
kernel's mean duration (CUDA events inside the library), achieved GB/s, clocks and power.

    python tools/perf_sweep.py [rows] [steps]
"""
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rag_serving_system_b200 import _cabi  # noqa: E402


class Power:
    def __init__(self):
        import pynvml
        pynvml.nvmlInit()
        self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(0)
        self.p, self.c, self.r = [], [], 0
        self._stop = threading.Event()

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def _run(self):
        while not self._stop.is_set():
            self.p.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000)
            self.c.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            self.r |= self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            time.sleep(0.01)

    def __exit__(self, *a):
        self._stop.set()
        self._t.join()


def run(ix, q_dev, B, k, metric, terms, steps, env, label, out):
    for key in ("RS_MMA_M", "RS_QRESIDENT", "RS_SUB", "RS_NBUF", "RS_CPS", "RS_INLINE_PREP", "RS_FUSED_TAIL", "RS_DEBUG_FLAGS", "RS_SHARED_BOUND", "RS_DOCM", "RS_LISTS_GLOBAL", "RS_DYN_SCHED", "RS_BUFC", "RS_LEAN"):
        os.environ.pop(key, None)
    os.environ.update({k_: str(v) for k_, v in env.items()})
    ix.reload_env_options()
    oi = torch.empty((B, k), dtype=torch.int64, device="cuda")
    os_ = torch.empty((B, k), dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    try:
        for _ in range(5):
            ix.search_device(q_dev.data_ptr(), B, k, oi.data_ptr(), os_.data_ptr(), metric, terms, st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with Power() as pw:
            e0.record()
            for _ in range(steps):
                ix.search_device(q_dev.data_ptr(), B, k, oi.data_ptr(), os_.data_ptr(), metric, terms, st)
            e1.record()
            torch.cuda.synchronize()
        stats = ix.stats()
        n = min(steps * stats["passes"], 256)
        kt = ix.kernel_times_ms(n)
        kms = float(kt.mean()) * stats["passes"]
        step_ms = e0.elapsed_time(e1) / steps
        rec = {
            "label": label, "B": B, "k": k, "metric": metric, "terms": terms, "env": env,
            "kernel_ms": round(kms, 4), "step_ms": round(step_ms, 4),
            "GBps": round(stats["corpus_bytes"] / (kms / 1e3) / 1e9, 1), "qps": round(B / (step_ms / 1e3), 1),
            "passes": stats["passes"], "stages": stats["stages"], "q_resident": stats["q_resident"],
            "sub": stats["sub"], "nbuf": stats["nbuf"], "mma_m": stats["mma_m"], "smem": stats["smem_bytes"],
            "sm_mhz": float(np.median(pw.c)) if pw.c else None, "power_w": float(np.median(pw.p)) if pw.p else None,
            "reasons": hex(pw.r), "kmin": round(float(kt.min()), 4), "kmax": round(float(kt.max()), 4),
        }
    except Exception as e:  # noqa
        rec = {"label": label, "env": env, "error": str(e)[:300]}
    print(json.dumps(rec), flush=True)
    out.append(rec)


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    which = sys.argv[3] if len(sys.argv) > 3 else "all"
    out = []
    D = 1024
    g = torch.Generator().manual_seed(0)
    if which in ("all", "bf16"):
        ix = _cabi.Index(rows, D, "bf16", max_batch=256, max_k=128)
        ix.fill_synthetic(1234, 0)
        ix.finalize()
        q = torch.randn(256, D, generator=g)
        q = (q / q.norm(dim=1, keepdim=True)).bfloat16().float().cuda()
        for env in ({"RS_MMA_M": 128}, {"RS_MMA_M": 64}, {"RS_MMA_M": 64, "RS_QRESIDENT": 0, "RS_SUB": 2, "RS_NBUF": 2},
                    {"RS_MMA_M": 64, "RS_QRESIDENT": 0, "RS_SUB": 4, "RS_NBUF": 1}, {"RS_MMA_M": 128}, {"RS_MMA_M": 64}):
            run(ix, q, 32, 10, "dot", 1, steps, env, "bf16 B32 k10", out)
        for B in (1, 8, 16, 64):
            run(ix, q, B, 10, "dot", 1, steps // 2, {}, f"bf16 B{B} k10", out)
        for env in ({}, {"RS_SUB": 2, "RS_NBUF": 2}):
            run(ix, q, 128, 10, "dot", 1, steps // 2, env, "bf16 B128 k10", out)
        run(ix, q, 256, 10, "dot", 1, steps // 4, {}, "bf16 B256 k10", out)
        run(ix, q, 32, 10, "cosine", 1, steps // 2, {}, "bf16 B32 k10 cosine", out)
        run(ix, q, 32, 10, "dot", 2, steps // 2, {}, "bf16 B32 k10 terms2", out)
        run(ix, q, 32, 100, "dot", 1, steps // 4, {}, "bf16 B32 k100", out)
        run(ix, q, 128, 100, "dot", 1, steps // 4, {}, "bf16 B128 k100", out)
        ix.close()
    if which in ("all", "fp32"):
        n3 = min(rows, 1_000_000)
        ix = _cabi.Index(n3, D, "fp32", max_batch=256, max_k=128, scale_exp=4)
        ix.fill_synthetic(1234, 0)
        ix.finalize()
        q = torch.randn(256, D, generator=g)
        q = (q / q.norm(dim=1, keepdim=True)).cuda()
        for B in (1, 32, 128):
            run(ix, q, B, 10, "dot", 2, steps, {}, f"fp32 1M B{B} k10", out)
        run(ix, q, 128, 10, "dot", 2, steps, {"RS_SUB": 1, "RS_NBUF": 2}, "fp32 1M B128 k10", out)
        run(ix, q, 32, 10, "cosine", 2, steps, {}, "fp32 1M B32 k10 cosine", out)
        ix.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "perf_sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
