"""Soak: many more random shapes than the committed property test, biased to the CTA-pair kernels' shapes (wide pair,
doc-major pair forced onto small corpora), each checked against the oracle like tests/test_gpu_parity.py does.

    python tools/soak.py [seed] [examples]        # round 2: seed 1, 400 examples, 0 failures
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.chdir(ROOT)
from oracle import retrieval_oracle as o
import test_gpu_parity as T
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
rows = T._PROP_ROWS
fails = 0
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 300):
    n = int(rng.choice([rng.integers(1, 700), rng.integers(700, len(rows))]))
    shape = rng.integers(0, 4)
    if shape == 0:   B, k = int(rng.integers(129, 257)), int(rng.integers(1, 129))          # wide pair
    elif shape == 1: B, k = int(rng.integers(17, 33)), int(rng.integers(1, 17))             # dm-pair, two accumulators
    elif shape == 2: B, k = int(rng.integers(33, 65)), int(rng.integers(1, 17))             # dm-pair, N = 64
    else:            B, k = int(rng.integers(1, 257)), int(rng.integers(1, 129))
    dtype = str(rng.choice(["bf16", "fp16", "fp32"])); metric = str(rng.choice(["dot", "cosine"]))
    two = bool(rng.integers(0, 2)) if shape != 2 else False
    try:
        T.test_property_any_shape_matches_oracle.hypothesis.inner_test(n, B, k, dtype, metric, two, int(rng.integers(0, 65536)), True)
    except Exception as e:  # noqa
        fails += 1
        print("FAIL", dict(n=n, B=B, k=k, dtype=dtype, metric=metric, two=two), repr(e)[:300], flush=True)
print("soak done, failures:", fails)
