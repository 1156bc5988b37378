"""GPU parity tests for spatial_autocorr: float64 scores within |d| <= 1e-5*|ref| + 1e-9 of the CPU restatement
(the tolerance BASELINE.json states; the near-zero atol is SURVEY.md H2).  In practice agreement is ~1e-13."""

from __future__ import annotations

import numpy as np
import pytest
import scipy.sparse as sp

import squidpy_b200 as sq
from oracle import ref
from squidpy_b200.gr import AutocorrPlan
from tools import synth

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-9
TIGHT = dict(rtol=1e-10, atol=1e-13)  # what float64 accumulation actually delivers


def _w(n_rows=40, n_cols=50, normalise=True):
    from sklearn.preprocessing import normalize

    g = synth.hex_graph(n_rows, n_cols)
    if normalise:
        normalize(g, norm="l1", axis=1, copy=False)
    return g


@pytest.mark.parametrize("mode", ["moran", "geary"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n_feat", [1, 31, 32, 70])
def test_dense_both_layouts(mode, dtype, n_feat):
    w = _w()
    n = w.shape[0]
    x = np.random.default_rng(n_feat).random((n, n_feat)).astype(dtype)  # obs x features
    exp = (ref.morans_i if mode == "moran" else ref.gearys_c)(w, x.T)
    plan = AutocorrPlan(w)
    plan.load(x, obs_major=True)
    got = plan.score(mode)
    np.testing.assert_allclose(got, exp, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(got, exp, **TIGHT)
    plan.load(np.ascontiguousarray(x.T), obs_major=False)
    np.testing.assert_array_equal(plan.score(mode), got)  # layouts are bit-identical (fixed reduction order)


@pytest.mark.parametrize("mode", ["moran", "geary"])
@pytest.mark.parametrize("fmt", ["csr", "csc"])
def test_sparse_layouts(mode, fmt):
    w = _w(50, 60, normalise=(mode == "moran"))
    n = w.shape[0]
    x = sp.random(n, 133, density=0.1, format=fmt, random_state=3, dtype=np.float32)  # obs x features like adata.X
    exp = (ref.morans_i if mode == "moran" else ref.gearys_c)(w, x.T.tocsr())
    plan = AutocorrPlan(w)
    plan.load(x, obs_major=True)
    got = plan.score(mode)
    np.testing.assert_allclose(got, exp, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(got, exp, **TIGHT)
    plan.load(x.toarray(), obs_major=True)
    np.testing.assert_array_equal(plan.score(mode), got)
    plan.load(x.T.tocsr(), obs_major=False)
    np.testing.assert_array_equal(plan.score(mode), got)


def test_constant_feature_nan_and_row_perm():
    w = _w()
    n = w.shape[0]
    x = np.random.default_rng(0).random((n, 5))
    x[:, 2] = 3.25
    plan = AutocorrPlan(w)
    plan.load(x, obs_major=True)
    got = plan.score("moran")
    assert np.isnan(got[2]) and np.isfinite(np.delete(got, 2)).all()
    idx = np.random.default_rng(1).permutation(n)
    x2 = np.delete(x, 2, axis=1)
    plan.load(x2, obs_major=True)
    for mode, f in (("moran", ref.morans_i), ("geary", ref.gearys_c)):
        np.testing.assert_allclose(plan.score(mode, row_perm=idx), f(w[idx, :], x2.T), **TIGHT)  # g[idx_shuffle, :], _ppatterns.py:271-272
    with pytest.raises(ValueError, match="not a permutation"):
        plan.score("moran", row_perm=np.zeros(n, np.int64))


def test_midsize_sparse_20k():
    g = _w(100, 200)
    co = synth.hex_coords(100, 200)
    x = synth.expression_csr(20000, 600, density=0.1, coords=co, seed=4)
    plan = AutocorrPlan(g)
    plan.load(x, obs_major=True)
    got = plan.score("moran")
    exp = ref.morans_i(g, x.T.tocsr())
    assert np.abs(got - exp).max() <= (RTOL * np.abs(exp) + ATOL).min()
    np.testing.assert_allclose(got, exp, **TIGHT)
    assert np.nanmax(exp) > 0.3 and np.isfinite(exp).sum() > 500  # the smooth genes are autocorrelated, none is constant
    np.testing.assert_allclose(plan.score("geary"), ref.gearys_c(g, x.T.tocsr()), **TIGHT)


@pytest.mark.parametrize("mode", ["moran", "geary"])
def test_api_matches_reference_golden(golden_dummy, dummy_adata, mode):
    """Full DataFrame (scores, analytic + permutation p-values, FDR, sort order) against the reference driver's
    output on dummy_adata (reference tests/graph/test_ppatterns.py:18-53 pins keys/columns)."""
    df = sq.gr.spatial_autocorr(dummy_adata, mode=mode, n_perms=10, seed=3, copy=True)
    cols = list(golden_dummy[f"autocorr_{mode}_columns"])
    assert list(df.columns) == cols and len(cols) == 9
    assert list(df.index) == list(golden_dummy[f"autocorr_{mode}_index"])
    np.testing.assert_allclose(df.to_numpy(dtype=np.float64), golden_dummy[f"autocorr_{mode}_values"], rtol=1e-7, atol=1e-12)
    sq.gr.spatial_autocorr(dummy_adata, mode=mode, transformation=False, two_tailed=True, genes=["g3", "g1", "g7"])
    key = "moranI" if mode == "moran" else "gearyC"
    out = dummy_adata.uns[key]
    assert list(out.index) == list(golden_dummy[f"autocorr_{mode}_nt_index"]) and out.shape[1] == 4
    np.testing.assert_allclose(out.to_numpy(dtype=np.float64), golden_dummy[f"autocorr_{mode}_nt_values"], rtol=1e-7, atol=1e-12)


def test_api_obs_and_obsm(golden_dummy, dummy_adata):
    df = sq.gr.spatial_autocorr(dummy_adata, attr="obs", genes=["cont"], copy=True)
    np.testing.assert_allclose(df.to_numpy(dtype=np.float64), golden_dummy["autocorr_obs_values"], rtol=1e-7, atol=1e-12)
    dummy_adata.obsm["emb"] = golden_dummy["X"][:, :4].copy()
    df = sq.gr.spatial_autocorr(dummy_adata, attr="obsm", layer="emb", copy=True)
    assert list(df.index.sort_values()) == [0, 1, 2, 3] and np.isfinite(df["I"]).all()
