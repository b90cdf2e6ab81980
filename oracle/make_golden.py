"""Generate the golden vectors in tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

It imports ``rag_service.core.retriever`` from /root/reference with ``cupy`` stubbed (the module
does ``import cupy`` at retriever.py:9 and CuPy is not installable here), runs the reference's
``SimpleRetriever`` (retriever.py:57-166) on the reference's own corpora, and stores inputs and
outputs.  The corpora themselves (data/*.npy, data/*_contexts.json) are copied as fixtures
because /root/reference does not exist on the GPU box.  TEST INFRASTRUCTURE ONLY.
"""
import json
import os
import shutil
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def import_reference():
    sys.modules.setdefault("cupy", types.ModuleType("cupy"))
    sys.path.insert(0, REF)
    import rag_service.core.retriever as ref  # noqa
    return ref


def main():
    ref = import_reference()
    os.makedirs(OUT, exist_ok=True)
    meta = {"numpy": np.__version__, "reference_class": "SimpleRetriever (retriever.py:57-166)",
            "cases": []}
    for name, B, k in (("short_facts", 32, 5), ("squad", 32, 10)):
        emb_src = os.path.join(REF, "data", f"{name}_embeddings.npy")
        ctx_src = os.path.join(REF, "data", f"{name}_contexts.json")
        shutil.copyfile(emb_src, os.path.join(OUT, f"{name}_embeddings.npy"))
        shutil.copyfile(ctx_src, os.path.join(OUT, f"{name}_contexts.json"))
        E = np.load(emb_src)
        ctx = json.load(open(ctx_src))
        N, D = E.shape
        rng = np.random.default_rng(0)
        src = rng.integers(0, N, size=B)
        # queries = seeded perturbations of corpus rows (no e5 model here)
        Q = (E[src] + 0.05 * rng.standard_normal((B, D))).astype(np.float32)
        ks_var = [int(x) for x in rng.integers(-1, 2 * k, size=B)]

        # index goldens: documents = stringified indices, so returned strings are the indices
        r_idx = ref.SimpleRetriever(E, [str(i) for i in range(N)])
        fixed = r_idx.batch_retrieve(Q, [k] * B)
        var = r_idx.batch_retrieve(Q, ks_var)
        self_top1 = [int(r_idx.retrieve(E[i], 1)[0]) for i in range(N)]
        # string goldens through the real contexts
        r_txt = ref.SimpleRetriever(E, ctx)
        txt = r_txt.batch_retrieve(Q, [k] * B)

        np.savez_compressed(
            os.path.join(OUT, f"{name}_golden.npz"),
            queries=Q, src_rows=src, k=k,
            idx_fixed=np.array([[int(s) for s in row] for row in fixed], dtype=np.int64),
            ks_var=np.array(ks_var, dtype=np.int64),
            idx_var=np.array([[int(s) for s in row] + [-1] * (2 * k - len(row)) for row in var],
                             dtype=np.int64),
            self_top1=np.array(self_top1, dtype=np.int64))
        json.dump(txt, open(os.path.join(OUT, f"{name}_golden_texts.json"), "w"))
        meta["cases"].append({"name": name, "N": int(N), "D": int(D), "B": B, "k": k})
    json.dump(meta, open(os.path.join(OUT, "META.json"), "w"), indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
