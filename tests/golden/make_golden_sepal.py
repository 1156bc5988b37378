"""Generates tests/golden/sepal.npz with the UNMODIFIED reference code (``/root/reference/src/squidpy/gr/_sepal.py``:
``_compute_idxs`` and ``_diffusion_genes`` / the numba kernel ``_diffusion``) on small lattices.  Build container only.

    python tests/golden/make_golden_sepal.py
"""

from __future__ import annotations

import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import _refload  # noqa: E402
from tools import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def square_graph(rows, cols):
    import scipy.sparse as sp

    idx = np.arange(rows * cols).reshape(rows, cols)
    src = np.r_[idx[:, :-1].ravel(), idx[:, 1:].ravel(), idx[:-1, :].ravel(), idx[1:, :].ravel()]
    dst = np.r_[idx[:, 1:].ravel(), idx[:, :-1].ravel(), idx[1:, :].ravel(), idx[:-1, :].ravel()]
    g = sp.csr_matrix((np.ones(src.size, np.float32), (src, dst)), shape=(rows * cols,) * 2)
    g.sort_indices()
    co = np.stack(np.meshgrid(np.arange(cols, dtype=float), np.arange(rows, dtype=float)), -1).reshape(-1, 2)
    return g, co


def make_case(name):
    if name == "hex":
        g, co = synth.hex_graph(20, 22), synth.hex_coords(20, 22)
        k = 6
    else:
        g, co = square_graph(18, 21)
        k = 4
    rng = np.random.default_rng(len(name))
    xy = (co - co.min(0)) / np.ptp(co, axis=0)
    cols = []
    for q in range(14):
        bump = np.exp(-((xy[:, 0] - rng.random()) ** 2 + (xy[:, 1] - rng.random()) ** 2) / (0.01 + 0.03 * q)) * (1 + q)
        noise = rng.random(len(xy)) * 0.25 * (q % 5)
        cols.append(bump + noise)
    cols.append(rng.poisson(0.3, len(xy)).astype(float))  # sparse counts
    cols.append(np.zeros(len(xy)))  # empty gene: entropy 0 from the start
    return g, co, k, np.stack(cols, axis=1)


def main():
    _refload.load()
    sp = importlib.import_module("squidpy.gr._sepal")
    out = {"meta": np.array("reference squidpy @ /root/reference (be17fcf6) gr/_sepal.py: _compute_idxs + _diffusion_genes (numba fastmath)")}
    for name in ("hex", "square"):
        g, co, k, vals = make_case(name)
        sat, sat_idx, unsat, unsat_idx = sp._compute_idxs(g, co, k, "l1")
        score = sp._diffusion_genes(vals, k == 6, 30000, sat, sat_idx, unsat, unsat_idx, 0.001, 1e-8, n_jobs=1, show_progress_bar=False)
        short = sp._diffusion_genes(vals, k == 6, 300, sat, sat_idx, unsat, unsat_idx, 0.001, 1e-8, n_jobs=1, show_progress_bar=False)
        out[f"{name}_sat"], out[f"{name}_sat_idx"], out[f"{name}_unsat"], out[f"{name}_unsat_idx"] = sat, sat_idx, unsat, unsat_idx
        out[f"{name}_score"], out[f"{name}_score_300"] = score, short
        print(name, np.round(score, 3), np.isnan(short).sum())
    np.savez_compressed(os.path.join(OUT, "sepal.npz"), **out)


if __name__ == "__main__":
    main()
