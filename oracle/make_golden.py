"""Generate the golden vectors in tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

It imports ``rag_service.core.retriever`` from /root/reference with ``cupy`` replaced by the NumPy
stand-in ``oracle/cupy_numpy_shim.py`` (the module does ``import cupy`` at retriever.py:9 and CuPy is
not installable here), runs the reference's ``SimpleRetriever`` (retriever.py:57-166, raw dot) AND
its live ``CupyRetriever`` (retriever.py:564-799, cosine; every ``cp.*`` call lands in NumPy) on the
reference's own corpora, and stores inputs and outputs.  The corpora themselves (data/*.npy, data/*_contexts.json) are copied as fixtures
because /root/reference does not exist on the GPU box.  TEST INFRASTRUCTURE ONLY.
"""
import json
import os
import shutil
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def import_reference():
    sys.path.insert(0, HERE)
    import cupy_numpy_shim
    cupy_numpy_shim.install()
    sys.path.insert(0, REF)
    import rag_service.core.retriever as ref  # noqa
    return ref


def main():
    ref = import_reference()
    os.makedirs(OUT, exist_ok=True)
    meta = {"numpy": np.__version__, "reference_class": "SimpleRetriever (retriever.py:57-166)",
            "reference_class_cosine": "CupyRetriever (retriever.py:564-799) with oracle/cupy_numpy_shim.py as cupy",
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

        # ---- cosine goldens: the live class, unmodified, NumPy standing in for CuPy.  Queries are
        # scaled per row (service queries are not normalised, batch_processor.py:91) so that cosine
        # scores differ from dot scores.
        import contextlib
        import io
        scale = rng.uniform(0.25, 4.0, size=(B, 1)).astype(np.float32)
        Qc = (Q * scale).astype(np.float32)
        c_idx = ref.CupyRetriever(E, [str(i) for i in range(N)])
        with contextlib.redirect_stdout(io.StringIO()):          # retriever.py:598 prints every call
            cfixed = c_idx.batch_retrieve(Qc, [k] * B)
            cvar = c_idx.batch_retrieve(Qc, ks_var)
            csingle = [c_idx.retrieve(Qc[i], k) for i in range(B)]
            S = np.asarray(c_idx.distance_cosine(Qc, E))         # the reference's own fp32 cosine matrix
            ctxt = ref.CupyRetriever(E, ctx).batch_retrieve(Qc, [k] * B)
        cidx_fixed = np.array([[int(s) for s in row] for row in cfixed], dtype=np.int64)
        assert S.dtype == np.float32 and S.shape == (B, N)
        assert cfixed == csingle, "retrieve() and batch_retrieve() of the reference disagree"
        top_scores = np.take_along_axis(S, cidx_fixed, axis=1)
        # goldens must be tie-free at fp32 so CuPy's unspecified tie order cannot matter
        srt = -np.sort(-S, axis=1)[:, : 2 * k + 1]
        assert (np.diff(srt, axis=1) < 0).all(), "exact fp32 tie among the top scores: pick another seed"
        np.savez_compressed(
            os.path.join(OUT, f"{name}_golden_cosine.npz"),
            queries=Qc, k=k, idx_fixed=cidx_fixed, scores_fixed=top_scores.astype(np.float32),
            ks_var=np.array(ks_var, dtype=np.int64),
            idx_var=np.array([[int(s) for s in row] + [-1] * (2 * k - len(row)) for row in cvar], dtype=np.int64))
        json.dump(ctxt, open(os.path.join(OUT, f"{name}_golden_cosine_texts.json"), "w"))
        meta["cases"].append({"name": name, "N": int(N), "D": int(D), "B": B, "k": k})
    json.dump(meta, open(os.path.join(OUT, "META.json"), "w"), indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
