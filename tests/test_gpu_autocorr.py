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
    plan.load(x.T.tocsr(), obs_major=False)
    np.testing.assert_array_equal(plan.score(mode), got)  # both sparse layouts end in the same ordered device layout
    xs = x.tocsr()
    perm = np.random.default_rng(0).permutation(xs.nnz)  # the same entries, shuffled inside every row (non-canonical order)
    rows = np.repeat(np.arange(n), np.diff(xs.indptr))[perm]
    o = np.argsort(rows, kind="stable")
    shuffled = sp.csr_matrix((xs.data[perm][o], xs.indices[perm][o], xs.indptr), shape=xs.shape)
    from squidpy_b200._lib import check, load

    out = np.empty(133)
    xp, xi = shuffled.indptr.astype(np.int64), shuffled.indices.astype(np.int32)
    check(load().sqb_autocorr_csr(plan._h, 0 if mode == "moran" else 1, xp.ctypes.data, xi.ctypes.data, shuffled.data.ctypes.data, 0, 1, 133, None, out.ctypes.data))
    np.testing.assert_array_equal(out, got)  # entry order of the input does not matter
    plan.load(x.toarray(), obs_major=True)
    np.testing.assert_allclose(plan.score(mode), got, **TIGHT)  # dense tiles: another formulation, same value to ~1e-14


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


@pytest.mark.parametrize("mode", ["moran", "geary"])
@pytest.mark.parametrize("kind", ["deg12_f32", "deg40_f32", "deg6_f64", "deg12_f64"])
def test_sparse_weight_formats(mode, kind):
    """Packed 8-lane rows (<= 7 entries, float32), packed 16-lane rows (<= 15), CSR rows (anything else / float64 weights)."""
    rng = np.random.default_rng(7)
    n, deg = 3000, int(kind.split("_")[0][3:])
    cols = np.stack([rng.choice(n, deg, replace=False) for _ in range(n)])
    dt = np.float32 if kind.endswith("f32") else np.float64
    w = sp.csr_matrix((rng.random(n * deg).astype(dt), cols.ravel(), np.arange(0, n * deg + 1, deg)), shape=(n, n))
    w.sort_indices()
    x = sp.random(n, 40, density=0.15, format="csr", random_state=1, dtype=np.float64)
    x.data = np.round(x.data * 8) / 4 - 0.5  # negative values and explicit zeros among the stored entries
    exp = (ref.morans_i if mode == "moran" else ref.gearys_c)(w, x.T.tocsr())
    plan = AutocorrPlan(w)
    plan.load(x, obs_major=True)
    np.testing.assert_allclose(plan.score(mode), exp, **TIGHT)
    idx = rng.permutation(n)
    exp_p = (ref.morans_i if mode == "moran" else ref.gearys_c)(w[idx, :], x.T.tocsr())
    np.testing.assert_allclose(plan.score(mode, row_perm=idx), exp_p, **TIGHT)
    plan.load(x.astype(np.float32), obs_major=True)
    np.testing.assert_allclose(plan.score(mode), (ref.morans_i if mode == "moran" else ref.gearys_c)(w, x.astype(np.float32).T.tocsr()), **TIGHT)


def test_sparse_run_to_run_bitwise_and_perm_batch():
    """float64 CSR by observation: the device transposition is ordered, so repeated loads give bit-identical scores;
    score_perms == one score(row_perm=...) per permutation == the oracle on g[idx, :]."""
    w = _w(60, 70)
    n = w.shape[0]
    x = sp.random(n, 90, density=0.2, format="csr", random_state=11, dtype=np.float64)
    plan = AutocorrPlan(w)
    runs = []
    for _ in range(3):
        plan.load(x, obs_major=True)
        runs.append((plan.score("moran"), plan.score("geary")))
    for a, b in runs[1:]:
        np.testing.assert_array_equal(a, runs[0][0])
        np.testing.assert_array_equal(b, runs[0][1])
    rng = np.random.default_rng(2)
    idx = np.stack([rng.permutation(n) for _ in range(5)])
    for mode, f in (("moran", ref.morans_i), ("geary", ref.gearys_c)):
        got = plan.score_perms(mode, idx)
        for p in range(5):
            np.testing.assert_array_equal(got[p], plan.score(mode, row_perm=idx[p]))
            np.testing.assert_allclose(got[p], f(w[idx[p], :], x.T.tocsr()), **TIGHT)
    bad = idx.copy()
    bad[3, 10] = bad[3, 11]
    with pytest.raises(ValueError, match="not a permutation"):
        plan.score_perms("moran", bad)
    plan.load(x.toarray(), obs_major=True)  # dense path through the same entry point
    np.testing.assert_allclose(plan.score_perms("moran", idx[:2])[1], ref.morans_i(w[idx[1], :], x.T.tocsr()), **TIGHT)


def test_sparse_invalid_input_is_rejected():
    from squidpy_b200._lib import check, load

    w = _w(20, 20)
    n = w.shape[0]
    plan = AutocorrPlan(w)
    lib = load()
    out = np.empty(3)

    def run(xp, xi, xv, layout, nf=3):
        xp, xi, xv = np.asarray(xp, np.int64), np.asarray(xi, np.int32), np.asarray(xv, np.float32)
        check(lib.sqb_autocorr_csr(plan._h, 0, xp.ctypes.data, xi.ctypes.data, xv.ctypes.data, 0, layout, nf, None, out.ctypes.data))

    run([0, 2, 2, 3], [5, 7, 1], [1, 2, 3], 0)  # fine (feature 1 is empty -> NaN)
    assert np.isnan(out[1]) and np.isfinite(out[[0, 2]]).all()
    with pytest.raises(ValueError, match="stored twice"):
        run([0, 2, 2, 3], [5, 5, 1], [1, 2, 3], 0)
    with pytest.raises(ValueError, match="out of range"):
        run([0, 2, 2, 3], [5, n, 1], [1, 2, 3], 0)
    with pytest.raises(ValueError, match="indptr"):
        run([0, 2, 1, 3], [5, 6, 1], [1, 2, 3], 0)
    xp = np.zeros(n + 1, np.int64)
    xp[4:] = 2
    with pytest.raises(ValueError, match="out of range"):
        run(xp, [0, 3], [1, 2], 1)  # observation 3 stores feature 3 of 3
    run(xp, [0, 2], [1, 2], 1)
    plan.close()


def test_sparse_million_observations_global_bitmap():
    """1M observations: the bitmap (8 B per 32 observations = 250 KB) no longer fits shared memory -> global scratch."""
    g = _w(1000, 1000)
    n = g.shape[0]
    x = sp.random(n, 6, density=0.05, format="csr", random_state=5, dtype=np.float32)
    plan = AutocorrPlan(g)
    plan.load(x, obs_major=True)
    np.testing.assert_allclose(plan.score("moran"), ref.morans_i(g, x.T.tocsr()), **TIGHT)
    np.testing.assert_allclose(plan.score("geary"), ref.gearys_c(g, x.T.tocsr()), **TIGHT)


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


@pytest.mark.parametrize("fmt", ["csr", "csc", "dense"])
def test_feature_shard_equals_slice_of_the_full_run(fmt):
    """Multi-GPU sharding loads a contiguous feature range per rank (`cols=`): for CSR-by-observation input the slice is cut
    while the matrix is staged (`sqb_autocorr_load_csr_cols`); the scores must be the corresponding slice of the full run, bit
    for bit, and an unsorted row must not lose entries."""
    w = _w(45, 50)
    n = w.shape[0]
    x = sp.random(n, 210, density=0.12, format="csr", random_state=8, dtype=np.float32)
    plan = AutocorrPlan(w)
    plan.load(x, obs_major=True)
    full_i, full_c = plan.score("moran"), plan.score("geary")
    m = {"csr": x, "csc": x.tocsc(), "dense": x.toarray()}[fmt]
    for lo, hi in ((0, 70), (70, 141), (141, 210), (5, 6)):
        plan.load(m, obs_major=True, cols=(lo, hi))
        assert plan.n_features == hi - lo
        got_i, got_c = plan.score("moran"), plan.score("geary")
        if fmt == "dense":
            np.testing.assert_allclose(got_i, full_i[lo:hi], **TIGHT)
            np.testing.assert_allclose(got_c, full_c[lo:hi], **TIGHT)
        else:
            np.testing.assert_array_equal(got_i, full_i[lo:hi])
            np.testing.assert_array_equal(got_c, full_c[lo:hi])
    # rows with descending column order: the wrapper sorts a copy before slicing
    rev = sp.csr_matrix((x.data.copy(), x.indices.copy(), x.indptr.copy()), shape=x.shape)
    for r in range(n):
        a, b = rev.indptr[r], rev.indptr[r + 1]
        rev.indices[a:b] = rev.indices[a:b][::-1]
        rev.data[a:b] = rev.data[a:b][::-1]
    rev.has_sorted_indices = False
    plan.load(rev, obs_major=True, cols=(70, 141))
    np.testing.assert_array_equal(plan.score("moran"), full_i[70:141])
