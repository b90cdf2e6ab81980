"""A/B of environment knobs on one box: same index, interleaved repeats, and BOTH power regimes.

Per variant and repeat: idle 1 s, then a short burst (cold clocks) is timed, then the loop keeps
running for `soak` seconds and the last stretch is timed again (power-capped clocks).

    python tools/ab_env.py ROWS B K DTYPE '[{}, {"RS_FUSED_TAIL": 0}, ...]' [reps] [soak_s] [metric]
"""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from rag_serving_system_b200 import _cabi
from perf_sweep import Power

KEYS = tuple(_cabi.OPTION_DEFAULTS)
rows, B, k = (int(x) for x in sys.argv[1:4])
dtype = sys.argv[4]
envs = json.loads(sys.argv[5])
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
soak = float(sys.argv[7]) if len(sys.argv) > 7 else 1.5
metric = sys.argv[8] if len(sys.argv) > 8 else "dot"
ix = _cabi.Index(rows, 1024, dtype, max_batch=256, max_k=128, scale_exp=4 if dtype == "fp32" else 0)
ix.fill_synthetic(1234, 0); ix.finalize()
g = torch.Generator().manual_seed(0)
q = torch.randn(256, 1024, generator=g); q = q / q.norm(dim=1, keepdim=True)
q = ({"bf16": q.bfloat16(), "fp16": q.half(), "fp32": q}[dtype]).float().cuda()
terms = 2 if dtype == "fp32" else 1
oi = torch.empty((256, 128), dtype=torch.int64, device="cuda")
os_ = torch.empty((256, 128), dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
B0, k0 = B, k


def go(n):
    for _ in range(n):
        ix.search_device(q.data_ptr(), B, k, oi.data_ptr(), os_.data_ptr(), metric, terms, st)


def timed(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); go(n); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


res = {}
for rep in range(reps):
    for env in envs:
        for key in KEYS:
            os.environ.pop(key, None)
        os.environ.update({k_: str(v) for k_, v in env.items() if not k_.startswith("_")})
        ix.reload_env_options()
        B, k = int(env.get("_B", B0)), int(env.get("_k", k0))      # "_B" / "_k" override the batch / k per variant
        name = json.dumps(env)
        try:
            go(3); torch.cuda.synchronize()
            time.sleep(1.0)
            go(5); torch.cuda.synchronize()
            est = timed(10)
            nb = max(8, int(0.05 / (est / 1e3)))            # ~50 ms burst
            burst = timed(nb)
            with Power() as pw:
                t0 = time.time()
                while time.time() - t0 < soak:
                    go(nb); torch.cuda.synchronize()
                hot = timed(2 * nb)
            stt = ix.stats()
            kt = float(ix.kernel_times_ms(min(2 * nb * stt["passes"], 256)).mean()) * stt["passes"]
            rec = {"burst_ms": burst, "hot_ms": hot, "hot_kernel_ms": kt, "mhz": float(np.median(pw.c[-20:])),
                   "w": float(np.median(pw.p[-20:])), "launches": stt["launches"], "passes": stt["passes"]}
        except Exception as e:  # noqa
            rec = {"error": str(e)[:200]}
        res.setdefault(name, []).append(rec)
        print(json.dumps({"env": env, "rep": rep, **{k_: (round(v, 4) if isinstance(v, float) else v) for k_, v in rec.items()}}), flush=True)
print("---- summary (median over reps)  rows=%d B=%d k=%d %s %s" % (rows, B, k, dtype, metric))
for name, rs in res.items():
    ok = [r for r in rs if "error" not in r]
    if not ok:
        print(name, "ERROR", rs[0].get("error")); continue
    med = lambda key: float(np.median([r[key] for r in ok]))
    gb = ix.stats()["corpus_bytes"] / 1e9
    print(f"{name:60s} burst {med('burst_ms'):.4f} ms ({gb / med('burst_ms') * ok[0]['passes']:.0f} GB/s/pass-eq)  hot {med('hot_ms'):.4f} ms  "
          f"hot-kernel {med('hot_kernel_ms'):.4f}  {med('mhz'):.0f} MHz {med('w'):.0f} W  launches {ok[0]['launches']}")
