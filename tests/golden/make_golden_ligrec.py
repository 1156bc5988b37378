"""Generates tests/golden/ligrec.npz with the UNMODIFIED reference code of the ligrec permutation test
(``/root/reference/src/squidpy/gr/_ligrec.py``: ``_analysis`` + ``_score_permutations``), loaded through the stub-import
loader.  ``numba_progress`` is not installed, so the reference's numba kernel is run through its own ``py_func`` (the same
source, interpreted) with a no-op progress object — small sizes only.  Only runnable in the build container.

    python tests/golden/make_golden_ligrec.py
"""

from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import _refload  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


class _Progress:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def update(self, n):
        pass


def load_ligrec():
    _refload.load()
    m = types.ModuleType("numba_progress")
    m.ProgressBar = _Progress
    sys.modules["numba_progress"] = m
    lg = importlib.import_module("squidpy.gr._ligrec")
    lg._score_permutations = lg._score_permutations.py_func
    lg.List = list
    lg.numba_threads = lambda n: contextlib.nullcontext()
    return lg


def make_case(seed, n_cells, n_genes, n_cls, kind):
    rng = np.random.default_rng(seed)
    if kind == "counts":
        x = rng.poisson(0.7, size=(n_cells, n_genes)).astype(np.float64)
    else:
        x = np.log1p(rng.gamma(0.6, 2.0, size=(n_cells, n_genes))) * (rng.random((n_cells, n_genes)) < 0.5)
    cl = rng.integers(0, n_cls, n_cells)
    cl[:n_cls] = np.arange(n_cls)
    inter = np.stack([rng.integers(0, n_genes, 3 * n_genes), rng.integers(0, n_genes, 3 * n_genes)], axis=1)
    inter = np.unique(inter, axis=0)
    cpairs = np.array([(a, b) for a in range(n_cls) for b in range(n_cls)], dtype=np.int64)
    return x, cl, inter, cpairs


def frame(x, cl, n_cls):
    df = pd.DataFrame(x, columns=list(range(x.shape[1])))
    df["clusters"] = pd.Categorical(cl, categories=list(range(n_cls)))
    return df


CASES = {"counts": (1, 240, 12, 4, "counts", 0.1, 40, 7), "lognorm": (2, 301, 17, 5, "lognorm", 0.5, 33, 11)}


def main():
    lg = load_ligrec()
    out = {"meta": np.array("reference squidpy @ /root/reference (be17fcf6) gr/_ligrec.py::_analysis with _score_permutations.py_func")}
    for name, (seed, n_cells, n_genes, n_cls, kind, thr, n_perms, pseed) in CASES.items():
        x, cl, inter, cpairs = make_case(seed, n_cells, n_genes, n_cls, kind)
        res = lg._analysis(frame(x, cl, n_cls), inter, cpairs, threshold=thr, n_perms=n_perms, seed=pseed, n_jobs=1, show_progress_bar=False)
        out[f"{name}_means"], out[f"{name}_pvalues"] = res.means, res.pvalues
        print(name, res.means.shape, float(np.nanmean(res.pvalues)), int(np.isnan(res.pvalues).sum()))
    np.savez_compressed(os.path.join(OUT, "ligrec.npz"), **out)


if __name__ == "__main__":
    main()
