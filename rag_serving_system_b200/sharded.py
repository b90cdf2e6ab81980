"""Row-sharded search across GPUs, one process per GPU (north_star (c)).

Each rank owns a contiguous row shard behind one ``_cabi.Index``.  A search is

    local fused score+select  ->  B x k (score, global index) candidates per rank
    exchange of the candidates over NVLink (B*k*12 bytes per rank), either
      exchange="nccl": ONE all-gather of the packed candidates (north_star (c)), or
      exchange="p2p" : the local merge kernel itself stores each row into every GPU's buffer
                       through CUDA-IPC peer pointers and raises an epoch flag (no collective call)
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

    def search_async(self, q_dev: torch.Tensor, k: int, metric: str = "dot", query_terms: int = 1):
        """Pipelined form for back-to-back searches (p2p exchange): the launch pushes this search's rows to
        the peers and never waits for them; the final merge of search i runs in the tail of search i+1 (or
        in ``flush``).  The returned tensors are therefore valid, in stream order, after the NEXT
        ``search_async`` / ``flush``; three output sets rotate, so keep at most two searches un-flushed
        before reading.  Falls back to ``search`` when there is nothing to pipeline."""
        if self.xchg is None:
            return self.search(q_dev, k, metric, query_terms)
        assert q_dev.is_cuda and q_dev.dtype == torch.float32 and q_dev.is_contiguous()
        B = q_dev.shape[0]
        key = ("async", B, k)
        if key not in self._bufs:
            dev = torch.device("cuda", self.device)
            self._bufs[key] = [0] + [(torch.empty((B, k), dtype=torch.int64, device=dev),
                                      torch.empty((B, k), dtype=torch.float32, device=dev)) for _ in range(3)]
        ring = self._bufs[key]
        ring[0] = (ring[0] + 1) % 3
        out_i, out_s = ring[1 + ring[0]]
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self.xchg.search_async(self.index, q_dev.data_ptr(), B, k, out_i.data_ptr(), out_s.data_ptr(), metric,
                               query_terms, stream)
        return out_i, out_s

    def flush(self):
        """Final merge of the last ``search_async`` (no-op when nothing is pending)."""
        if self.xchg is not None:
            self.xchg.flush(self.index, torch.cuda.current_stream(self.device).cuda_stream)

    def search_host(self, q_host: torch.Tensor, k: int, metric: str = "dot", query_terms: int = 1):
        """End-to-end call with HOST buffers: pinned H2D of the queries, search, D2H of the result into
        pinned host tensors (reused by the next call with the same B and k), one synchronisation."""
        dev = torch.device("cuda", self.device)
        B = q_host.shape[0]
        key = ("host", B, k)
        if key not in self._bufs:
            self._bufs[key] = (torch.empty((B, q_host.shape[1]), dtype=torch.float32, device=dev),
                               torch.empty((B, k), dtype=torch.int64).pin_memory(),
                               torch.empty((B, k), dtype=torch.float32).pin_memory())
        q_dev, pin_i, pin_s = self._bufs[key]
        q_dev.copy_(q_host, non_blocking=True)
        oi, os_ = self.search(q_dev, k, metric, query_terms)
        pin_i.copy_(oi, non_blocking=True)
        pin_s.copy_(os_, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return pin_i, pin_s

    def close(self):
        if self.xchg is not None:
            self.flush()
            torch.cuda.synchronize(self.device)
            if dist.is_initialized():
                dist.barrier(group=self.group)      # nobody unmaps a buffer a peer may still write
            self.xchg.close()
            self.xchg = None
        self.index.close()
