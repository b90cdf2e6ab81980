#!/usr/bin/env python
"""Headline benchmark: queries/sec (and HBM GB/s vs the measured peak) of the fused similarity +
top-k search, B=32, k=10, over a synthetic unit-normalised 10M x 1024 bf16 corpus row-sharded over
N B200s (BASELINE.json configs[3]).

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path over one batch: query prep -> fused score+select over the
rank's shard -> candidate merge (+ one NCCL all-gather and the final on-device merge when N>1).
`value` is timed on the device with inputs resident in HBM; `e2e` is the same metric through the
host-buffer C-ABI call (pinned H2D of the queries and D2H of the result inside the timed region).
`--impl reference` times the reference's CPU path (oracle port of SimpleRetriever,
retriever.py:57-166) on the host cores for the same config.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "queries/sec"
SEED = 1234


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16", "fp32"])
    ap.add_argument("--metric", default="dot", choices=["dot", "cosine"])
    ap.add_argument("--cpu-sample-rows", type=int, default=200_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default=None, choices=["nccl", "p2p"], help="multi-GPU candidate exchange")
    ap.add_argument("--deferred-exchange", action="store_true",
                    help="p2p exchange: pipeline the final merge of search i into the tail of search i+1 "
                         "(ShardedSearcher.search_async + flush) instead of waiting for the peers inside every search; "
                         "measured equal at 2 and 8 GPUs, so the default stays the synchronous form")
    return ap.parse_args()


def workload_config(a, world):
    return {
        "workload": f"synthetic unit-norm {a.rows}x{a.dim} {a.dtype} corpus, B={a.batch}, k={a.k}, {a.metric}; "
                    f"BASELINE.json configs[3]",
        "rows": a.rows, "dim": a.dim, "batch": a.batch, "k": a.k, "similarity": a.metric,
        "sharding": f"row-sharded x{world}", "cache": "inputs larger than L2 (no flush needed)",
        "exchange": ((a.exchange or os.environ.get("RS_EXCHANGE", "p2p")) +
                     (" (deferred: final merge of search i in the tail of search i+1, flush inside the timed region)"
                      if a.deferred_exchange and (a.exchange or os.environ.get("RS_EXCHANGE", "p2p")) == "p2p" else ""))
        if world > 1 else "none",
    }


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvml), runs during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    # NVML clocks-event-reason bits
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, device_index):
        self.samples, self.power, self.mask, self.max_mhz, self.limit_w = [], [], 0, None, None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            try:
                self.limit_w = pynvml.nvmlDeviceGetEnforcedPowerLimit(self.h) / 1000.0
            except Exception:
                pass
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                self.power.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        reasons = sorted(name for bit, name in self.REASONS.items() if self.mask & bit)
        out = {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
               "reasons": reasons, "reasons_mask": hex(self.mask), "samples": len(self.samples)}
        if self.power:
            out["power_w_max"] = round(max(self.power), 1)
            out["power_limit_w"] = self.limit_w
        return out


# ------------------------------------------------------------------------------------------------
# CPU baseline: the reference's NumPy path (port), bounded sample, linear extrapolation stated
# ------------------------------------------------------------------------------------------------
def time_cpu_port(rows_f32, Q, k, full_rows, reps=3):
    from oracle.retrieval_oracle import SimpleRetrieverPort          # the checker, timed as baseline only
    try:
        from threadpoolctl import threadpool_info
        threads = max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        threads = os.cpu_count() or 1
    r = SimpleRetrieverPort(rows_f32, range(rows_f32.shape[0]))
    ks = [k] * Q.shape[0]
    r.batch_retrieve(Q, ks)                                           # warm-up
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        r.batch_retrieve(Q, ks)
        ts.append(time.perf_counter() - t)
    med = float(np.median(ts))
    qps_sample = Q.shape[0] / med
    qps_full = qps_sample * rows_f32.shape[0] / full_rows            # one GEMV per query: linear in N
    return {
        "value": qps_full, "unit": METRIC, "cores": int(threads), "kind": "port",
        "sample": f"SimpleRetriever port (retriever.py:57-166) on the first {rows_f32.shape[0]} rows as fp32, "
                  f"B={Q.shape[0]}, k={k}: median {med:.3f} s/batch = {qps_sample:.2f} q/s on the sample, scaled "
                  f"linearly to {full_rows} rows; numpy {np.__version__}, {os.cpu_count()} host cpus",
    }


def run_reference(a):
    """Reference arm: the reference's own CPU implementation of the path (oracle port) on host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import retrieval_oracle as o
    n = min(a.cpu_sample_rows, a.rows)
    rows = o.round_to_dtype(o.synth_rows(SEED, np.arange(n), a.dim), o.DTYPE_NAMES[a.dtype])
    Q, _ = o.synth_queries(5, n, a.batch, a.dim, corpus_seed=SEED, dtype=o.DTYPE_NAMES[a.dtype])
    from oracle.retrieval_oracle import SimpleRetrieverPort
    try:
        from threadpoolctl import threadpool_info
        threads = max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        threads = os.cpu_count() or 1
    r = SimpleRetrieverPort(rows, range(n))
    ks = [a.k] * a.batch
    for _ in range(max(1, min(a.warmup, 2))):
        r.batch_retrieve(Q, ks)
    steps = max(1, min(a.steps, 5))
    t = time.perf_counter()
    for _ in range(steps):
        r.batch_retrieve(Q, ks)
    el = (time.perf_counter() - t) / steps
    qps = a.batch / el * n / a.rows
    base = {
        "value": qps, "unit": METRIC, "cores": int(threads), "kind": "port",
        "sample": f"SimpleRetriever port (retriever.py:57-166) on the first {n} rows as fp32, B={a.batch}, k={a.k}: "
                  f"{el:.3f} s/batch = {a.batch / el:.2f} q/s on the sample, scaled linearly to {a.rows} rows; "
                  f"numpy {np.__version__}, {os.cpu_count()} host cpus",
    }
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": METRIC, "n_gpus": a.gpus,
        "steps": steps, "warmup": a.warmup, "ms_per_step": el * 1e3 * a.rows / n, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(a, a.gpus), "cpu_baseline": base,
        "e2e": {"value": qps, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
        return

    import torch
    import torch.distributed as dist
    from rag_serving_system_b200 import _cabi
    from rag_serving_system_b200.sharded import ShardedSearcher

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: no CUDA device, and there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if a.gpus != world and rank == 0:
        print(f"warning: --gpus {a.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)

    B, k = a.batch, a.k
    s = ShardedSearcher(a.rows, a.dim, a.dtype, max_batch=max(B, 64), max_k=max(k, 16), device=local_rank,
                        exchange=a.exchange)
    s.fill_synthetic(SEED)

    # queries: perturbed corpus rows (rank 0's shard), representable in the storage dtype so one
    # tensor-core term is exact; identical on every rank
    if rank == 0:
        g = torch.Generator().manual_seed(5)
        ids = torch.randint(0, s.index.n_rows, (B,), generator=g)
        base = np.stack([s.index.get_rows(int(i), 1)[0] for i in ids])
        q = torch.from_numpy(base) + 0.01 * torch.randn(B, a.dim, generator=g)
        q = q / q.norm(dim=1, keepdim=True)
        q = q.to({"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[a.dtype]).float()
    else:
        ids = torch.zeros(B, dtype=torch.int64)
        q = torch.empty(B, a.dim)
    q_dev = q.to(dev)
    ids_dev = ids.to(dev)
    if world > 1:
        dist.broadcast(q_dev, 0)
        dist.broadcast(ids_dev, 0)
    q_host = q_dev.cpu().pin_memory()
    terms = 2 if a.dtype == "fp32" else 1

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing
    for _ in range(max(a.warmup, 3)):
        oi, os_ = s.search(q_dev, k, a.metric, terms)
    sync_all()
    assert bool((oi[:, 0] == ids_dev).all()), "top-1 of a perturbed corpus row must be that row"
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        sync_all()
        ev0.record()
        if world > 1 and s.xchg is not None and a.deferred_exchange:
            for _ in range(a.steps):                      # back-to-back searches: deferred final merge
                oi, os_ = s.search_async(q_dev, k, a.metric, terms)
            s.flush()                                     # the last search's final merge, inside the timed region
        else:
            for _ in range(a.steps):
                s.search(q_dev, k, a.metric, terms)
        ev1.record()
        sync_all()
    assert bool((oi[:, 0] == ids_dev).all()), "timed loop: top-1 of a perturbed corpus row must be that row"
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    kt = s.index.kernel_times_ms(min(a.steps, 256))
    kms = torch.tensor([float(kt.mean())], device=dev)
    kms_ranks = [float(kt.mean())]
    if world > 1:
        gathered_k = [torch.zeros_like(kms) for _ in range(world)]
        dist.all_gather(gathered_k, kms)
        kms_ranks = [round(float(t.item()), 4) for t in gathered_k]
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    qps = B * a.steps / (total_ms / 1e3)
    st = s.index.stats()

    # ---------------- end-to-end through the host-buffer API
    q_np = q_host.numpy()
    e2e_steps = a.steps

    def e2e_call():
        if world == 1:
            return s.index.search(q_np, k, a.metric, terms)        # C ABI rs_search: host in, host out
        return s.search_host(q_host, k, a.metric, terms)

    for _ in range(3):
        e2e_call()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_call()
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_qps = B * e2e_steps / float(e2e_s.item())

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        shard_bytes = st["corpus_bytes"]                 # algorithmic bytes: the shard, streamed once per batch
        achieved = shard_bytes / (float(kms.item()) / 1e3) / 1e9
        traffic = None
        prof = os.path.join(ROOT, "profiles", "fused_traffic.json")
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof)).get(f"{a.rows}x{a.dim}_{a.dtype}_g{world}")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": qps, "unit": METRIC, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16" if a.dtype == "bf16" else ("f16" if a.dtype == "fp16" else "f16x2 (fp32 split)"),
            "data": "synthetic", "config": workload_config(a, world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "fused_search_kernel",
                         "kernel_ms": float(kms.item()), "algorithmic_bytes_per_launch": shard_bytes,
                         "frac_of_8TBps_spec": achieved / 8000.0, "kernel_ms_per_rank": kms_ranks},
            "e2e": {"value": e2e_qps, "unit": METRIC, "h2d_bytes_per_step": B * a.dim * 4,
                    "d2h_bytes_per_step": B * k * 12, "api": "rs_search (C ABI, host buffers)" if world == 1
                    else "ShardedSearcher.search_host"},
            "gpu_launches": st["launches"] * a.steps + (a.steps if (world > 1 and s.xchg is None) else 0),
            "clocks": clk.summary(),
            "kernel_stats": {"grid": st["grid"], "stages": st["stages"], "smem_bytes": st["smem_bytes"],
                             "passes": st["passes"], "query_terms": st["query_terms"]},
        }
        if world == 1 and not a.no_cpu_baseline:
            n = min(a.cpu_sample_rows, s.index.n_rows)
            rows = s.index.get_rows(0, n)
            line["cpu_baseline"] = time_cpu_port(rows, q_np, k, a.rows)
        print(json.dumps(line))
    s.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
