"""GPU parity tests of the ligrec permutation test: p-values and means identical to the reference's
``_analysis`` on golden inputs (tests/golden/ligrec.npz, produced by the unmodified reference code), permutation counts equal
to the numpy restatement at larger sizes, and the public ``ligrec`` frames."""

from __future__ import annotations

import os

import numpy as np
import pandas as pd
import pytest

import squidpy_b200 as sq
from oracle import ref
from squidpy_b200.gr import ligrec_analysis
from tests.golden.make_golden_ligrec import CASES, frame, make_case

pytestmark = pytest.mark.gpu
GOLD = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "ligrec.npz"), allow_pickle=False))


@pytest.mark.parametrize("name", list(CASES))
def test_analysis_matches_reference_golden(name):
    seed, n_cells, n_genes, n_cls, kind, thr, n_perms, pseed = CASES[name]
    x, cl, inter, cpairs = make_case(seed, n_cells, n_genes, n_cls, kind)
    res = ligrec_analysis(frame(x, cl, n_cls), inter, cpairs, threshold=thr, n_perms=n_perms, seed=pseed)
    np.testing.assert_array_equal(res.pvalues, GOLD[f"{name}_pvalues"])  # counts / n_perms: exact, NaN where not tested
    np.testing.assert_array_equal(res.means, GOLD[f"{name}_means"])
    assert np.isnan(GOLD[f"{name}_pvalues"]).any() or name == "counts"


def test_counts_vs_numpy_restatement_midsize():
    rng = np.random.default_rng(3)
    n_cells, n_genes, n_cls, P = 20000, 150, 12, 64
    x = np.log1p(rng.gamma(0.5, 2.0, (n_cells, n_genes))) * (rng.random((n_cells, n_genes)) < 0.3)
    cl = rng.integers(0, n_cls, n_cells)
    inter = np.unique(np.stack([rng.integers(0, n_genes, 400), rng.integers(0, n_genes, 400)], 1), axis=0)
    cpairs = np.array([(a, b) for a in range(n_cls) for b in range(n_cls)])
    df = frame(x, cl, n_cls)
    res = ligrec_analysis(df, inter, cpairs, threshold=0.05, n_perms=P, seed=5)
    g = df.groupby("clusters", observed=True)
    mean_obs = g.mean().values
    inv = 1.0 / np.maximum(g.size().values.astype(np.float64), 1)
    valid = ~np.isnan(res.pvalues)
    counts = ref.ligrec_counts(x, cl, n_cls, ref.spawn_states(5, P), inv, mean_obs, inter, cpairs, valid)
    np.testing.assert_array_equal(np.where(valid, res.pvalues, 0.0), counts / P)
    assert valid.mean() > 0.5 and 0.2 < np.nanmean(res.pvalues) < 0.8


def test_public_api_frames():
    rng = np.random.default_rng(0)
    n, genes = 500, ["A", "B", "C", "D", "E", "F"]
    X = rng.poisson(1.0, (n, len(genes))).astype(np.float32)
    obs = pd.DataFrame({"cl": pd.Categorical(rng.choice(["x", "y", "z"], n))}, index=[f"c{i}" for i in range(n)])
    ad = sq.AnnDataLite(X=X, obs=obs, var=pd.DataFrame(index=genes))
    inter = pd.DataFrame({"source": ["a", "b", "A_C", "zz", "d"], "target": ["b", "c", "d", "a", "E_F"], "note": list("vwxyz")})
    res = sq.gr.ligrec(ad, "cl", interactions=inter, use_raw=False, n_perms=50, seed=1, copy=True, threshold=0.0)
    assert set(res) == {"means", "pvalues", "metadata"}
    assert res["means"].shape == res["pvalues"].shape == (4, 9)  # the interaction with the unknown gene 'ZZ' is dropped
    assert list(res["means"].index.names) == ["source", "target"] and res["means"].columns.nlevels == 2
    assert ("x", "y") in res["means"].columns and ("A", "B") in res["means"].index
    pv = res["pvalues"].to_numpy(dtype=float)
    assert np.nanmin(pv) >= 0.0 and np.nanmax(pv) <= 1.0
    again = sq.gr.ligrec(ad, "cl", interactions=inter, use_raw=False, n_perms=50, seed=1, copy=True, threshold=0.0)
    np.testing.assert_array_equal(again["pvalues"].to_numpy(dtype=float), pv)
    sq.gr.ligrec(ad, "cl", interactions=[("A", "B"), ("C", "D")], use_raw=False, n_perms=20, seed=2, corr_method="fdr_bh", clusters=["x", "y"])
    out = ad.uns["cl_ligrec"]
    assert out["pvalues"].shape == (2, 4)
    with pytest.raises(ValueError, match="Invalid cluster"):
        sq.gr.ligrec(ad, "cl", interactions=[("A", "B")], use_raw=False, clusters=["x", "nope"])
    with pytest.raises(NotImplementedError):
        sq.gr.ligrec(ad, "cl", use_raw=False)
    with pytest.raises(AttributeError, match="raw"):
        sq.gr.ligrec(ad, "cl", interactions=[("A", "B")])
