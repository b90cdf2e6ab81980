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
    for dtype, metric, k, B, exch in (("bf16", "dot", 10, 32, "p2p"), ("bf16", "dot", 10, 32, "nccl"),
                                      ("fp32", "cosine", 5, 7, "p2p"), ("fp16", "dot", 100, 3, "nccl"),
                                      ("bf16", "dot", 17, 100, "p2p")):
        s = ShardedSearcher(N, D, dtype, max_batch=128, max_k=128, scale_exp=4 if dtype == "fp32" else 0,
                            exchange=exch)
        s.fill_synthetic(1234)
        Q, src = o.synth_queries(9, N, B, D, dtype=o.DTYPE_NAMES[dtype])
        q_dev = torch.from_numpy(Q).cuda()
        terms = 2 if dtype == "fp32" else 1
        for _ in range(3):                      # repeated calls exercise both parities of the exchange buffer
            idx, sc = s.search(q_dev, k, metric, terms)
        torch.cuda.synchronize()
        idx, sc = idx.cpu().numpy(), sc.cpu().numpy()
        hidx, hsc = s.search_host(torch.from_numpy(Q).pin_memory(), k, metric, terms)
        assert np.array_equal(hidx.numpy(), idx) and np.array_equal(hsc.numpy(), sc)
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
        if exch == "p2p":
            # pipelined (deferred) exchange: the final merge of search i runs in the tail of search i+1 or in
            # flush(); every result must equal the synchronous search of the same queries, also when B and k
            # change between calls and when synchronous searches are interleaved
            seq = [(Q, k), (Q[::-1].copy(), k), (Q[: max(1, B // 2)].copy(), max(1, k // 2)), (Q, k), (Q[::-1].copy(), k)]
            want = []
            for Qi, ki in seq:
                wi, ws = s.search(torch.from_numpy(Qi).cuda(), ki, metric, terms)
                want.append((wi.clone(), ws.clone()))
            got, prev = [], None
            for Qi, ki in seq:
                cur = s.search_async(torch.from_numpy(Qi).cuda(), ki, metric, terms)
                if prev is not None:
                    got.append((prev[0].clone(), prev[1].clone()))      # valid after the NEXT call (stream order)
                prev = cur
            s.flush()
            got.append((prev[0].clone(), prev[1].clone()))
            mid_i, mid_s = s.search_async(q_dev, k, metric, terms)          # pending search, then a synchronous one
            si, ss = s.search(q_dev, k, metric, terms)
            torch.cuda.synchronize()
            assert torch.equal(mid_i, si) and torch.equal(mid_s, ss)
            for n, ((gi, gs), (wi, ws)) in enumerate(zip(got, want)):
                assert torch.equal(gi, wi) and torch.equal(gs, ws), f"deferred exchange: search {n} differs"
            if rank == 0:
                print(f"deferred exchange ok world={world} {dtype} {metric} k={k} B={B}", flush=True)
        s.close()
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
