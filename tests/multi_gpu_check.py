"""Run under torchrun (one rank per GPU): row-sharded search == oracle on the whole corpus.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/multi_gpu_check.py

Launched by tests/test_gpu_multi.py; exits non-zero on any mismatch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import retrieval_oracle as o  # noqa: E402
from rag_serving_system_b200.sharded import ShardedSearcher  # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    N, D = 100_003, 1024                       # uneven shards
    ok = True
    # (dtype, metric, k, B, exchange, RS_DMPAIR): the last two run the doc-major CTA-pair kernel (forced onto these small
    # shards) with the exchange in its tail - fp32 storage (two accumulators) and 32 < B <= 64 (N = 64)
    for dtype, metric, k, B, exch, dmpair in (("bf16", "dot", 10, 32, "p2p", 1), ("bf16", "dot", 10, 32, "nccl", 1),
                                              ("fp32", "cosine", 5, 7, "p2p", 1), ("fp16", "dot", 100, 3, "nccl", 1),
                                              ("bf16", "dot", 17, 100, "p2p", 1), ("fp16", "dot", 100, 256, "p2p", 1),
                                              ("fp16", "cosine", 100, 256, "p2p", 1), ("fp32", "cosine", 10, 32, "p2p", 2),
                                              ("bf16", "dot", 16, 64, "p2p", 2)):
        os.environ["RS_DMPAIR"] = str(dmpair)       # read once, by rs_create
        s = ShardedSearcher(N, D, dtype, max_batch=256, max_k=128, scale_exp=4 if dtype == "fp32" else 0,
                            exchange=exch)
        s.fill_synthetic(1234)
        Q, src = o.synth_queries(9, N, B, D, dtype=o.DTYPE_NAMES[dtype])
        q_dev = torch.from_numpy(Q).cuda()
        terms = 2 if dtype == "fp32" else 1
        for _ in range(3):                      # repeated calls exercise both parities of the exchange buffer
            idx, sc = s.search(q_dev, k, metric, terms)
        torch.cuda.synchronize()
        s.index.check()
        assert (s.index.stats()["mma_m"] == 256) == (dmpair == 2 or B > 128), s.index.stats()   # the CTA-pair kernels ran where meant
        idx, sc = idx.cpu().numpy(), sc.cpu().numpy()
        # host-buffer call (C ABI, CUDA-graph replay from the 2nd call on); only rank 0 asks for the result
        for _ in range(3):
            hidx, hsc = s.search_host(Q, k, metric, terms, want_result=(rank == 0 or exch == "nccl"))
        if rank == 0 or exch == "nccl":
            assert np.array_equal(hidx, idx) and np.array_equal(hsc, sc)
        # every rank must hold the same answer
        gathered = [None] * world
        dist.all_gather_object(gathered, (idx.tobytes(), sc.tobytes()))
        assert all(g == gathered[0] for g in gathered), "ranks disagree"
        if rank == 0:
            rows = o.synth_rows(1234, np.arange(N), D)
            S = o.scores(Q.astype(np.float64), o.stored_values(rows, o.DTYPE_NAMES[dtype], 4 if dtype == "fp32" else 0),
                         metric, accum=np.float64)
            sw = o.check_topk(idx, sc, S, k, rtol=1e-5 if dtype == "fp32" else 2e-3, atol=1e-6)
            assert (idx[:, 0] == src).all()
            print(f"multi-gpu ok world={world} {dtype} {metric} k={k} B={B} exchange={exch} near-tie swaps={sw}", flush=True)
        dist.barrier()
        if exch == "p2p":
            # overlapped launches (RS_OVERLAP, programmatic dependent launch): the scan of search i+1 runs while the
            # exchange tail of search i drains; every result must equal the serial search of the same queries, also
            # when B and k change between calls
            seq = [(Q, k), (Q[::-1].copy(), k), (Q[: max(1, B // 2)].copy(), max(1, k // 2)), (Q, k), (Q[::-1].copy(), k)] * 2
            qs = [torch.from_numpy(Qi).cuda() for Qi, _ in seq]
            want = []
            for qd, (Qi, ki) in zip(qs, seq):
                wi, ws = s.search(qd, ki, metric, terms)
                want.append((wi.clone(), ws.clone()))
            torch.cuda.synchronize()
            s.set_overlap(True)
            got = []
            for qd, (Qi, ki) in zip(qs, seq):
                gi, gs = s.search(qd, ki, metric, terms)
                got.append((gi.clone(), gs.clone()))      # clone: a kernel of torch's between two searches is allowed
            torch.cuda.synchronize()
            s.index.check()
            s.set_overlap(False)
            for n, ((gi, gs), (wi, ws)) in enumerate(zip(got, want)):
                assert torch.equal(gi, wi) and torch.equal(gs, ws), f"overlapped launches: search {n} differs"
            if rank == 0:
                print(f"overlapped launches ok world={world} {dtype} {metric} k={k} B={B}", flush=True)
        s.close()
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
