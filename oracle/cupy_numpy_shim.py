"""A NumPy-backed stand-in for the ``cupy`` module.  TEST INFRASTRUCTURE ONLY.

The reference's live class ``CupyRetriever`` (retriever.py:564-799) holds no arithmetic of its own:
every operation is a CuPy call (``cp.asarray``, ``@``, ``cp.linalg.norm``, ``cp.clip``,
``cp.argpartition``, ``cp.argsort``, ``cp.empty``, ``.get()``; call sites retriever.py:582,596-609,
652,662-672,738,746,757-773).  CuPy (un-pinned ``cupy-cuda12x``, Dockerfile:12; hint 13.4.1 at
requirements.txt:84) cannot be installed here or on the GPU box, so ``oracle/make_golden.py`` runs
the UNMODIFIED reference class with this module registered as ``sys.modules["cupy"]``.  The NumPy
function behind each name follows CuPy's published semantics for that call:

* ``asarray`` / ``empty`` / ``clip`` / ``argsort`` / ``linalg.norm`` / ``@``: same definition in
  both libraries (CuPy mirrors the NumPy API); arithmetic stays fp32 because the reference coerces
  its inputs to fp32 first (retriever.py:576-578,736-737).
* ``argpartition``: CuPy documents that it performs a FULL sort along the axis ("currently
  implemented as a full sort", cupy/_sorting/search.py in 13.x), i.e. ``argpartition(a, kth)`` ==
  ``argsort(a)``.  Restated here as a stable argsort.  The tie order of CuPy's thrust sort is not
  specified; goldens are generated on tie-free inputs (checked by the generator).
* ``ndarray.get()``: device -> host copy; here the identity on a NumPy array.

Only ``oracle/make_golden.py`` and ``tests/`` import this module.
"""
from __future__ import annotations

import types

import numpy as np


class ndarray(np.ndarray):
    """np.ndarray plus the one CuPy-only method the reference calls."""

    def get(self):
        return np.asarray(self)


def _wrap(a):
    return np.asarray(a).view(ndarray)


def asarray(a, dtype=None):
    return _wrap(np.asarray(a, dtype=dtype))


def empty(shape, dtype=float):
    return _wrap(np.empty(shape, dtype=dtype))


def clip(a, a_min, a_max):
    return _wrap(np.clip(np.asarray(a), a_min, a_max))


def argsort(a, axis=-1):
    return _wrap(np.argsort(np.asarray(a), axis=axis, kind="stable"))


def argpartition(a, kth, axis=-1):
    # CuPy: "currently implemented as a full sort" -> identical to argsort
    return _wrap(np.argsort(np.asarray(a), axis=axis, kind="stable"))


def _norm(x, ord=None, axis=None, keepdims=False):
    return _wrap(np.linalg.norm(np.asarray(x), ord=ord, axis=axis, keepdims=keepdims))


linalg = types.SimpleNamespace(norm=_norm)
int64 = np.int64
float32 = np.float32


def install():
    """Register this module as ``cupy`` (before importing rag_service.core.retriever)."""
    import sys
    sys.modules["cupy"] = sys.modules[__name__]
    return sys.modules[__name__]
