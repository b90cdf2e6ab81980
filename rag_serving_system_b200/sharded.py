"""Row-sharded search across GPUs, one process per GPU (north_star (c)).

Each rank owns a contiguous row shard behind one ``_cabi.Index``.  A search is

    local fused score+select  ->  B x k (score, global index) candidates per rank
    exchange of the candidates over NVLink (B*k*12 bytes per rank), either
      exchange="nccl": ONE all-gather of the packed candidates (north_star (c)), or
      exchange="p2p" : the tail of the scan kernel itself stores each row into every GPU's buffer through
                       CUDA-IPC peer pointers as 16-byte {index, score, epoch} entries (no collective call)
    on-device merge of the G sorted lists per query on every rank

The path has exactly one exchange step.  The reference has no multi-GPU search at all (it scales
by whole-corpus replicas, request_queue.py:49).

``gather_candidates`` is backend-agnostic (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _cabi


def shard_range(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row shards: rank g owns [g*ceil(N/G), min(N,(g+1)*ceil(N/G)))."""
    per = -(-n_rows // world)
    lo = min(n_rows, rank * per)
    return lo, min(n_rows, lo + per)


def packed_bytes(B: int, k: int) -> int:
    """One rank's candidate message: int64 idx[B][k] followed by float32 score[B][k], padded to 16 B
    so every rank's block stays 8-byte aligned inside the gathered buffer."""
    return (B * k * 12 + 15) // 16 * 16


def pack_views(buf: torch.Tensor, B: int, k: int):
    """(idx int64[B,k], score float32[B,k]) views into a packed uint8 buffer of packed_bytes(B,k)."""
    idx = buf[: B * k * 8].view(torch.int64).view(B, k)
    sc = buf[B * k * 8: B * k * 12].view(torch.float32).view(B, k)
    return idx, sc


def gather_candidates(local: torch.Tensor, out: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather every rank's packed candidate buffer into out[world, nbytes]."""
    dist.all_gather_into_tensor(out.view(-1), local.view(-1), group=group)
    return out


class ShardedSearcher:
    """Per-rank handle of a row-sharded corpus.  ``search`` returns the same global top-k on
    every rank (device tensors)."""

    def __init__(self, n_rows_total: int, dim: int, dtype: str = "bf16", *, max_batch: int = 256,
                 max_k: int = 128, device: Optional[int] = None, group=None, scale_exp: int = 0,
                 exchange: Optional[str] = None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = torch.cuda.current_device() if device is None else device
        self.n_total, self.dim, self.dtype = n_rows_total, dim, dtype
        self.row0, self.row1 = shard_range(n_rows_total, self.world, self.rank)
        if self.row1 <= self.row0:
            raise ValueError("more ranks than rows")
        self.index = _cabi.Index(self.row1 - self.row0, dim, dtype, device=self.device, row_offset=self.row0,
                                 max_batch=max_batch, max_k=max_k, scale_exp=scale_exp)
        self._bufs = {}
        self.exchange = exchange or os.environ.get("RS_EXCHANGE", "p2p")
        if self.exchange not in ("nccl", "p2p"):
            raise ValueError("exchange must be 'nccl' or 'p2p'")
        self.xchg = None
        if self.world > 1 and self.exchange == "p2p":
            self.xchg = _cabi.Exchange(self.device, self.world, self.rank, max_batch, max_k)
            handles = [None] * self.world
            dist.all_gather_object(handles, self.xchg.export_handle(), group=group)
            for g, h in enumerate(handles):
                self.xchg.import_handle(g, h)
            dist.barrier(group=group)

    # ---- ingest
    def fill_synthetic(self, seed: int):
        self.index.fill_synthetic(seed, self.row0)
        self.index.finalize()

    def load_rows(self, rows_global: np.ndarray):
        """rows_global: the FULL [N, D] host array (or anything sliceable); this rank ingests its shard."""
        step = 1 << 16
        for r in range(self.row0, self.row1, step):
            n = min(step, self.row1 - r)
            self.index.load_rows(np.asarray(rows_global[r:r + n]), r - self.row0)
        self.index.finalize()

    # ---- search
    def _buffers(self, B: int, k: int):
        key = (B, k)
        if key not in self._bufs:
            dev = torch.device("cuda", self.device)
            nb = packed_bytes(B, k)
            local = torch.empty(nb, dtype=torch.uint8, device=dev)
            gathered = torch.empty((self.world, nb), dtype=torch.uint8, device=dev)
            out_i = torch.empty((B, k), dtype=torch.int64, device=dev)
            out_s = torch.empty((B, k), dtype=torch.float32, device=dev)
            self._bufs[key] = (local, gathered, out_i, out_s)
        return self._bufs[key]

    def search(self, q_dev: torch.Tensor, k: int, metric: str = "dot", query_terms: int = 1):
        """q_dev: float32 [B, D] on this rank's GPU (every rank passes the same queries)."""
        assert q_dev.is_cuda and q_dev.dtype == torch.float32 and q_dev.is_contiguous()
        B = q_dev.shape[0]
        local, gathered, out_i, out_s = self._buffers(B, k)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        if self.world == 1:
            self.index.search_device(q_dev.data_ptr(), B, k, out_i.data_ptr(), out_s.data_ptr(), metric,
                                     query_terms, stream)
            return out_i, out_s
        if self.xchg is not None:
            self.xchg.search(self.index, q_dev.data_ptr(), B, k, out_i.data_ptr(), out_s.data_ptr(), metric,
                             query_terms, stream)
            return out_i, out_s
        li, ls = pack_views(local, B, k)
        self.index.search_device(q_dev.data_ptr(), B, k, li.data_ptr(), ls.data_ptr(), metric, query_terms, stream)
        gather_candidates(local, gathered, self.group)
        nb = packed_bytes(B, k)
        gi = gathered.data_ptr()
        _cabi.merge_device(self.device, gi + B * k * 8, gi, self.world, B, k, k, out_i.data_ptr(), out_s.data_ptr(),
                           stream, score_list_stride=nb // 4, idx_list_stride=nb // 8)
        return out_i, out_s

    def set_overlap(self, on: bool):
        """Let consecutive ``search`` calls on one stream overlap (RS_OVERLAP, programmatic dependent launch):
        the scan of search i+1 runs while the merge / exchange tail of search i drains.  The queries of a
        search must then not be produced by the kernel that immediately precedes it on the stream."""
        self.index.set_option("RS_OVERLAP", 1 if on else 0)

    def search_host(self, q_host, k: int, metric: str = "dot", query_terms: int = 1, want_result: bool = True):
        """End-to-end call with HOST buffers (a NumPy array or a CPU torch tensor): H2D of the queries from the
        library's pinned staging, search (+ exchange), D2H of the result, one synchronisation - all inside the C
        ABI (rs_search / rs_search_exchange_host), replayed from a CUDA graph.  ``want_result=False`` on ranks
        that only take part (no D2H).  The NCCL exchange keeps the torch-level sequence."""
        q_np = q_host.numpy() if isinstance(q_host, torch.Tensor) else q_host
        if self.world == 1:
            return self.index.search(q_np, k, metric, query_terms)
        if self.xchg is not None:
            return self.xchg.search_host(self.index, q_np, k, metric, query_terms, want_result)
        dev = torch.device("cuda", self.device)
        B = q_np.shape[0]
        key = ("host", B, k)
        if key not in self._bufs:
            self._bufs[key] = (torch.empty((B, q_np.shape[1]), dtype=torch.float32).pin_memory(),
                               torch.empty((B, q_np.shape[1]), dtype=torch.float32, device=dev),
                               torch.empty((B, k), dtype=torch.int64).pin_memory(),
                               torch.empty((B, k), dtype=torch.float32).pin_memory())
        q_pin, q_dev, pin_i, pin_s = self._bufs[key]
        q_pin.numpy()[...] = q_np
        q_dev.copy_(q_pin, non_blocking=True)
        oi, os_ = self.search(q_dev, k, metric, query_terms)
        pin_i.copy_(oi, non_blocking=True)
        pin_s.copy_(os_, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return pin_i.numpy(), pin_s.numpy()

    def close(self):
        if self.xchg is not None:
            torch.cuda.synchronize(self.device)
            if dist.is_initialized():
                dist.barrier(group=self.group)      # nobody unmaps a buffer a peer may still write
            self.xchg.close()
            self.xchg = None
        self.index.close()
