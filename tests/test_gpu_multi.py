"""Multi-GPU parity (needs >= 2 B200s on the box; skipped otherwise)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


def test_sharded_search_under_torchrun():
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("multi-gpu ok") == 5
    assert res.stdout.count("deferred exchange ok") == 3          # every p2p case also runs the pipelined form


def test_single_process_multi_device_class(squad):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    from oracle import retrieval_oracle as o
    from rag_serving_system_b200 import B200Retriever
    r = B200Retriever(squad.E, squad.contexts, dtype="fp32", metric="dot", devices=list(range(min(n, 4))))
    idx, sc = r.search(squad.Q, squad.k)
    assert (idx == squad.idx_fixed).all()
    assert r.batch_retrieve(squad.Q, [squad.k] * len(squad.Q)) == squad.texts
    r.close()
