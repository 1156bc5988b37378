"""GPU parity tests for co_occurrence (float32 pair counts) and Ripley's L (float64 pair counts): integers bit-exact."""

from __future__ import annotations

import numpy as np
import pandas as pd
import pytest

import squidpy_b200 as sq
from oracle import ref
from squidpy_b200.gr import cooc_counts, pair_counts
from tools import synth

pytestmark = pytest.mark.gpu


def test_cooc_golden_lattice_integers(golden_dummy):
    g = golden_dummy
    s32 = g["xy"].astype(np.float32)
    got = cooc_counts(s32[:, 0], s32[:, 1], g["cooc_interval"][1:] ** 2, g["cl"].astype(np.int32), 3)
    assert got.dtype == np.int64 and got.shape == (3, 3, 49)
    np.testing.assert_array_equal(got, g["cooc_counts"])
    np.testing.assert_array_equal(got[:, :, 0], [[138, 119, 110], [119, 114, 92], [110, 92, 98]])  # KAT-3


def test_cooc_jittered_fma_semantics(golden_pairs):
    g = golden_pairs
    thr = g["interval"][1:] ** 2
    got = cooc_counts(g["pts"][:, 0], g["pts"][:, 1], thr, g["labs"], 4)
    np.testing.assert_array_equal(got, g["cooc_counts"])  # == the reference's numba kernel, FMA contraction included
    nofma = cooc_counts(g["pts"][:, 0], g["pts"][:, 1], thr, g["labs"], 4, use_fma=False)
    np.testing.assert_array_equal(nofma, ref.occur_count(g["pts"][:, 0], g["pts"][:, 1], thr, g["labs"], 4, use_fma=False))


@pytest.mark.parametrize("n,k,L", [(1, 1, 3), (2, 2, 1), (257, 1, 5), (3000, 7, 49), (5000, 3, 200), (9000, 20, 49)])
def test_cooc_vs_oracle(n, k, L):
    rng = np.random.default_rng(n + k)
    pts = (rng.random((n, 2)) * 500).astype(np.float32)
    labs = rng.integers(0, k, n).astype(np.int32)
    if k > 2:
        labs[labs == 1] = 0  # an empty label
    thr = np.sort(rng.random(L).astype(np.float32) * 400.0) ** 2
    if L > 4:
        thr[3] = thr[2]  # duplicate radius
    got = cooc_counts(pts[:, 0], pts[:, 1], thr, labs, k)
    np.testing.assert_array_equal(got, ref.occur_count(pts[:, 0], pts[:, 1], thr, labs, k))
    np.testing.assert_array_equal(got, got.transpose(1, 0, 2))
    assert (np.diff(got, axis=2) >= 0).all()


def test_cooc_unsorted_thresholds_and_shards():
    rng = np.random.default_rng(0)
    pts = (rng.random((2500, 2)) * 100).astype(np.float32)
    labs = rng.integers(0, 4, 2500).astype(np.int32)
    thr = np.array([900.0, 4.0, 2500.0, 100.0], np.float32)
    exp = ref.occur_count(pts[:, 0], pts[:, 1], thr, labs, 4)
    np.testing.assert_array_equal(cooc_counts(pts[:, 0], pts[:, 1], thr, labs, 4), exp)
    parts = [cooc_counts(pts[:, 0], pts[:, 1], thr, labs, 4, shard=(r, 3)) for r in range(3)]
    np.testing.assert_array_equal(sum(parts), exp)  # tile shards add up (multi-GPU path)


def test_cooc_midsize_20k():
    rng = np.random.default_rng(4)
    n = 20000
    pts = (rng.random((n, 2)) * 2.0e4).astype(np.float32)
    labs = rng.integers(0, 20, n).astype(np.int32)
    iv = np.linspace(50.0, 1.4e4, 50, dtype=np.float32)
    got = cooc_counts(pts[:, 0], pts[:, 1], iv[1:] ** 2, labs, 20)
    np.testing.assert_array_equal(got, ref.occur_count(pts[:, 0], pts[:, 1], iv[1:] ** 2, labs, 20))


def test_pair_counts_golden(golden_dummy, golden_pairs):
    g = golden_dummy
    groups = [g["xy"][g["cl"] == c].astype(np.float64) for c in range(3)]
    got = pair_counts(groups, g["ripley_support"])
    np.testing.assert_array_equal(got, g["ripley_tp"])  # == sklearn KDTree.two_point_correlation
    np.testing.assert_array_equal(got[0][:6], [77, 79, 83, 93, 123, 157])  # KAT-4
    p = golden_pairs
    np.testing.assert_array_equal(pair_counts([p["P"]], p["support"])[0], p["tp"])
    gx, gy = np.meshgrid(np.arange(60), np.arange(60))
    lattice = np.stack([gx.ravel(), gy.ravel()], 1).astype(np.float64)
    np.testing.assert_array_equal(pair_counts([lattice], p["lattice_support"])[0], p["lattice_tp"])  # exact ties at every radius


def test_pair_counts_vs_oracle_groups_and_shards():
    rng = np.random.default_rng(1)
    groups = [rng.random((m, 2)) * 1000 for m in (1500, 0, 1, 3000, 700)]
    sup = np.linspace(0, 700, 50)
    exp = np.stack([ref.pair_counts(gp, sup) if len(gp) else np.zeros(50, np.int64) for gp in groups])
    np.testing.assert_array_equal(pair_counts(groups, sup), exp)
    parts = [pair_counts(groups, sup, shard=(r, 2)) for r in range(2)]
    np.testing.assert_array_equal(sum(parts), exp)


def test_api_co_occurrence_golden(golden_dummy, dummy_adata, golden_cfg1):
    occ, interval = sq.gr.co_occurrence(dummy_adata, "cluster", copy=True)
    assert occ.dtype == np.float64 and occ.shape == (3, 3, 49) and interval.dtype == np.float32
    np.testing.assert_array_equal(interval, golden_dummy["cooc_interval"])
    np.testing.assert_allclose(occ, golden_dummy["cooc_occ"], rtol=1e-12)
    sq.gr.co_occurrence(dummy_adata, "cluster", interval=np.array([300.0, 10.0, 50.0, 120.5]))
    d = dummy_adata.uns["cluster_co_occurrence"]
    np.testing.assert_array_equal(d["interval"], golden_dummy["cooc_interval_explicit"])
    np.testing.assert_allclose(d["occ"], golden_dummy["cooc_occ_explicit"], rtol=1e-12)
    # configs[0] grid (exact lattice ties everywhere)
    lab = pd.Series(pd.Categorical.from_codes(golden_cfg1["codes"].astype(int), categories=[f"c{i:03d}" for i in range(10)]))
    ad = synth.make_adata(synth.hex_coords(71, 71), None, lab)
    occ, interval = sq.gr.co_occurrence(ad, "cluster", interval=12, copy=True)
    np.testing.assert_array_equal(interval, golden_cfg1["cooc_interval"])
    np.testing.assert_allclose(occ, golden_cfg1["cooc_occ"], rtol=1e-12)


@pytest.mark.parametrize("mode", ["L", "F", "G"])
def test_api_ripley_golden(golden_dummy, dummy_adata, mode):
    res = sq.gr.ripley(dummy_adata, "cluster", mode=mode, n_simulations=20, n_observations=300, n_steps=50, seed=7, copy=True)
    # reference tests/graph/test_ripley.py:45-98 shapes
    assert res[f"{mode}_stat"].shape == (50 * 3, 3) and res["sims_stat"].shape == (50 * 20, 3)
    assert res["bins"].shape == (50,) and res["pvalues"].shape == (3, 50)
    np.testing.assert_allclose(res["bins"], golden_dummy[f"ripley_{mode}_bins"], rtol=1e-15)
    np.testing.assert_allclose(res[f"{mode}_stat"]["stats"].to_numpy(), golden_dummy[f"ripley_{mode}_stat"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(res["sims_stat"]["stats"].to_numpy(), golden_dummy[f"ripley_{mode}_sims"], rtol=1e-12, atol=1e-15)
    np.testing.assert_array_equal(res["pvalues"], golden_dummy[f"ripley_{mode}_pvalues"])
    assert res[f"{mode}_stat"]["stats"].iloc[0] == 0.0


def test_api_co_occurrence_unused_category_in_the_middle(dummy_adata):
    """the reference sizes its output by the PRESENT categories but addresses it with raw codes (undefined behaviour
    when a category in the middle is unused).  Here present codes are remapped: the result must equal the one obtained after
    dropping the unused category, and the `occ` axis has one entry per present category."""
    import pandas as pd

    ad = dummy_adata
    codes = np.asarray(ad.obs["cluster"].cat.codes).copy()
    codes[codes == 1] = 2  # category '1' stays declared but unused
    ad.obs["cluster"] = pd.Categorical.from_codes(codes, categories=["0", "1", "2"])
    occ, iv = sq.gr.co_occurrence(ad, "cluster", interval=12, copy=True)
    assert occ.shape == (2, 2, 11)
    ad.obs["cluster"] = ad.obs["cluster"].cat.remove_unused_categories()
    occ2, iv2 = sq.gr.co_occurrence(ad, "cluster", interval=12, copy=True)
    np.testing.assert_array_equal(occ, occ2)
    np.testing.assert_array_equal(iv, iv2)
