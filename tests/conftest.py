import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


class Golden:
    def __init__(self, name):
        self.name = name
        self.E = np.load(os.path.join(GOLD, f"{name}_embeddings.npy"))
        self.contexts = json.load(open(os.path.join(GOLD, f"{name}_contexts.json")))
        g = np.load(os.path.join(GOLD, f"{name}_golden.npz"))
        self.Q = g["queries"]
        self.k = int(g["k"])
        self.idx_fixed = g["idx_fixed"]
        self.ks_var = [int(x) for x in g["ks_var"]]
        self.idx_var = g["idx_var"]
        self.self_top1 = g["self_top1"]
        self.texts = json.load(open(os.path.join(GOLD, f"{name}_golden_texts.json")))
        # outputs of the unmodified reference CupyRetriever (cosine) with NumPy standing in for CuPy
        c = np.load(os.path.join(GOLD, f"{name}_golden_cosine.npz"))
        self.cos_Q = c["queries"]
        self.cos_idx_fixed = c["idx_fixed"]
        self.cos_scores_fixed = c["scores_fixed"]
        self.cos_idx_var = c["idx_var"]
        self.cos_texts = json.load(open(os.path.join(GOLD, f"{name}_golden_cosine_texts.json")))


@pytest.fixture(scope="session", params=["short_facts", "squad"])
def golden(request):
    return Golden(request.param)


@pytest.fixture(scope="session")
def squad():
    return Golden("squad")


@pytest.fixture(scope="session")
def short_facts():
    return Golden("short_facts")
