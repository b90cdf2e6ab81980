"""First-light diagnostics on a B200 (run under gpurun; writes gpurun_out/diag.log).

Not a test: a progressive log that shows WHICH stage breaks and HOW (layout, MMA descriptors,
selection, merge).  Uses numpy + ctypes only (no torch) so it starts fast.
"""
import json
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import retrieval_oracle as o  # noqa: E402
from rag_serving_system_b200 import _cabi  # noqa: E402

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "diag.log"), "a")
DT = {"bf16": o.DTYPE_BF16, "fp16": o.DTYPE_FP16, "fp32": o.DTYPE_FP32}


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    LOG.write(s + "\n")
    LOG.flush()


def case(name):
    def deco(fn):
        def run():
            log(f"=== {name}")
            t = time.time()
            try:
                fn()
                log(f"--- {name}: done in {time.time() - t:.2f}s")
            except Exception:
                log(f"!!! {name}: EXCEPTION\n" + traceback.format_exc())
        run.__name__ = name
        return run
    return deco


@case("layout_synth")
def layout_synth():
    for dt in ("bf16", "fp16", "fp32"):
        N, D = 300, 1024
        ix = _cabi.Index(N, D, dt, max_batch=64, max_k=16)
        ix.fill_synthetic(1234, 0)
        ix.finalize()
        rows = ix.get_rows(0, N)
        exp32 = o.synth_rows(1234, np.arange(N), D)
        exp = o.stored_values(exp32, DT[dt]).astype(np.float32)
        log(dt, "get_rows == oracle:", bool((rows == exp).all()), "maxdiff", float(np.abs(rows - exp).max()))
        img = ix.get_image(0, ix.stats()["corpus_bytes"]).view(np.uint16)
        eimg = o.tile_corpus(o.corpus_planes_u16(exp32, DT[dt]))
        log(dt, "image == oracle:", bool((img == eimg).all()), "nmismatch", int((img != eimg).sum()), img.size, eimg.size)
        nr = ix.get_norms(0, N)
        log(dt, "norm maxerr", float(np.abs(nr - np.linalg.norm(exp.astype(np.float64), axis=1)).max()))
        ix.close()


@case("layout_load")
def layout_load():
    E = np.load(os.path.join(ROOT, "tests/golden/squad_embeddings.npy"))
    for dt in ("bf16", "fp16", "fp32"):
        ix = _cabi.Index(E.shape[0], E.shape[1], dt, max_batch=64, max_k=16)
        ix.load_rows(E)
        ix.finalize()
        rows = ix.get_rows(0, E.shape[0])
        exp = o.stored_values(E, DT[dt]).astype(np.float32)
        log(dt, "load round trip exact:", bool((rows == exp).all()), float(np.abs(rows - exp).max()))
        ix.close()


def run_search(E, Q, k, dt, metric, terms, label, show=True):
    N, D = E.shape
    ix = _cabi.Index(N, D, dt, max_batch=max(256, Q.shape[0]), max_k=max(128, k))
    ix.load_rows(E)
    ix.finalize()
    t = time.time()
    idx, sc = ix.search(Q, k, metric, terms)
    el = time.time() - t
    Dv = o.stored_values(E, DT[dt])
    Qv = Q.astype(np.float64) if (terms == 2 or dt == "fp32") else o.round_to_dtype(Q, DT[dt]).astype(np.float64)
    S = o.scores(Qv, Dv, metric, accum=np.float64)
    eidx, esc = o.topk_rows(S, k)
    same = (idx == eidx)
    log(label, f"N={N} B={Q.shape[0]} k={k} {dt} {metric} terms={terms}: idx match {same.mean():.4f}",
        f"rows fully equal {int(same.all(axis=1).sum())}/{Q.shape[0]}",
        f"score maxrel {float(np.nanmax(np.abs(sc - esc) / np.maximum(np.abs(esc), 1e-30))):.3e}",
        f"t={el*1e3:.2f}ms stats={ix.stats()}")
    try:
        sw = o.check_topk(idx, sc, S, k, rtol=(1e-5 if dt == "fp32" else 2e-3), atol=1e-6)
        log(label, "check_topk OK, near-tie swaps:", sw)
    except AssertionError as e:
        log(label, "check_topk FAILED:", str(e)[:600])
        if show:
            log("  got idx[0]", idx[0].tolist())
            log("  exp idx[0]", eidx[0].tolist())
            log("  got sc[0] ", sc[0].tolist())
            log("  exp sc[0] ", esc[0].tolist())
    ix.close()
    return idx, sc


@case("onehot")
def onehot():
    # doc j = (1 + (j % 7)/8) * e_{(j*37) % D};  query i = e_{(i*37*5) % D}: exact scores, reveals K permutations
    N, D, B, k = 1000, 1024, 16, 4
    E = np.zeros((N, D), np.float32)
    for j in range(N):
        E[j, (j * 37) % D] = 1 + (j % 7) / 8
    Q = np.zeros((B, D), np.float32)
    for i in range(B):
        Q[i, (i * 37 * 5) % D] = 1
    run_search(E, Q, k, "bf16", "dot", 1, "onehot")


@case("squad_bf16")
def squad_bf16():
    E = np.load(os.path.join(ROOT, "tests/golden/squad_embeddings.npy"))
    g = np.load(os.path.join(ROOT, "tests/golden/squad_golden.npz"))
    Q = g["queries"]
    run_search(E, Q, 10, "bf16", "dot", 1, "squad")
    run_search(E, Q, 10, "bf16", "dot", 2, "squad")
    run_search(E, Q, 10, "fp16", "dot", 1, "squad")
    run_search(E, Q, 10, "fp16", "dot", 2, "squad")
    idx, sc = run_search(E, Q, 10, "fp32", "dot", 2, "squad")
    log("squad fp32 vs reference golden idx equal:", bool((idx == g["idx_fixed"]).all()))
    run_search(E, Q, 10, "bf16", "cosine", 1, "squad")
    run_search(E, Q, 10, "fp32", "cosine", 2, "squad")
    run_search(E, Q, 100, "bf16", "dot", 1, "squad-k100")
    run_search(E, Q[:1], 10, "bf16", "dot", 1, "squad-B1")
    run_search(E, np.concatenate([Q, Q[::-1] * 1.5, Q * 0.5, Q[:4]]), 5, "bf16", "dot", 1, "squad-B100")


@case("multi_tile")
def multi_tile():
    N, D = 128 * 148 * 3 + 77, 1024
    E = o.synth_rows(7, np.arange(N), D)
    rng = np.random.default_rng(1)
    ids = rng.integers(0, N, 32)
    Q = o.round_to_bf16(E[ids] + 0.02 * rng.standard_normal((32, D)).astype(np.float32))
    run_search(E, Q, 10, "bf16", "dot", 1, "multi")
    run_search(E, Q, 10, "fp32", "dot", 2, "multi")
    run_search(E, Q, 10, "bf16", "cosine", 1, "multi")


@case("bandwidth")
def bandwidth():
    import ctypes as C
    N, D, B, k = 2_000_000, 1024, 32, 10
    ix = _cabi.Index(N, D, "bf16", max_batch=64, max_k=16)
    t = time.time()
    ix.fill_synthetic(1234, 0)
    ix.finalize()
    log("fill 2M rows:", f"{time.time()-t:.2f}s")
    Q, ids = o.synth_queries(5, N, B, D, dtype=o.DTYPE_BF16)
    for it in range(5):
        t = time.time()
        idx, sc = ix.search(Q, k, "dot", 1)
        el = time.time() - t
        log(f"iter {it}: host-timed {el*1e3:.3f} ms -> {N*D*2/el/1e9:.1f} GB/s (includes H2D/D2H + 3 launches)")
    log("top1 == source row:", int((idx[:, 0] == ids).sum()), "/", B, "stats", ix.stats())
    ix.close()


if __name__ == "__main__":
    log("lib", _cabi.lib_path(), "version", _cabi.load().rs_version())
    want = sys.argv[1:]
    for fn in (layout_synth, layout_load, onehot, squad_bf16, multi_tile, bandwidth):
        if not want or fn.__name__ in want:
            fn()
    log("ALL DONE")
