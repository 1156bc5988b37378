"""Generates tests/golden/*.npz by running the UNMODIFIED reference code (``/root/reference/src/squidpy``) through the
stub-import loader ``oracle/_refload.py``.  Only runnable in the build container (the GPU box has no reference tree);
the outputs are committed.  Versions used: see the ``meta`` entry of each file.

    python tests/golden/make_golden.py
"""

from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import _refload, ref  # noqa: E402
from squidpy_b200._adata import AnnDataLite  # noqa: E402
from tools import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def _mt(p, alpha=0.05, method="fdr_bh"):
    assert method == "fdr_bh"
    return None, ref.multipletests_fdr_bh(p), None, None


def main():
    import numba
    import scipy
    import sklearn

    m = _refload.load(morans_i=ref.morans_i, gearys_c=ref.gearys_c, multipletests=_mt)
    nh, pp, rp, nb = m["nh"], m["pp"], m["rp"], m["nb"]
    meta = np.array(
        f"numpy {np.__version__}; numba {numba.__version__}; scipy {scipy.__version__}; sklearn {sklearn.__version__}; "
        f"reference squidpy @ /root/reference (be17fcf6); Moran/Geary values come from oracle/c/oracle.c (scanpy absent)"
    )

    # ------------------------------------------------------------------ dummy_adata recipe (tests/conftest.py:109-117)
    r = np.random.RandomState(100)
    X = r.rand(200, 100)
    cl = r.randint(0, 3, 200)
    xy = np.stack([r.randint(0, 500, 200), r.randint(0, 500, 200)], axis=1)
    adj, _ = nb.KNNBuilder(n_neighs=6).build(xy)
    cats = pd.Categorical.from_codes(cl, categories=["0", "1", "2"])
    libs = pd.Categorical(["A"] * 100 + ["B"] * 100)

    def mk():
        obs = pd.DataFrame({"cluster": cats, "library": libs, "cont": X[:, 0]}, index=[str(i) for i in range(200)])
        var = pd.DataFrame(index=[f"g{i}" for i in range(100)])
        return AnnDataLite(X=X.copy(), obs=obs, var=var, obsm={"spatial": xy.copy()}, obsp={"spatial_connectivities": adj.copy()})

    out = {"meta": meta, "X": X, "cl": cl, "xy": xy, "adj_indptr": adj.indptr, "adj_indices": adj.indices, "adj_data": adj.data}
    res = nh.nhood_enrichment(mk(), "cluster", n_perms=20, seed=42, copy=True, show_progress_bar=False)
    out["nhood_z"], out["nhood_count"] = res.zscore, res.counts
    res = nh.nhood_enrichment(mk(), "cluster", library_key="library", n_perms=20, seed=42, copy=True, show_progress_bar=False)
    out["nhood_lib_z"], out["nhood_lib_count"] = res.zscore, res.counts
    # per-permutation counts (helper level)
    from squidpy._utils import spawn_generators

    f3 = nh._create_function(3)
    ind, ptr = adj.indices.astype(np.uint32), adj.indptr.astype(np.uint32)
    out["nhood_perms"] = nh._nhood_enrichment_helper(list(range(20)), f3, ind, ptr, cl.astype(np.uint32), None, 3, spawn_generators(42, 20)).astype(np.uint32)
    out["nhood_lib_perms"] = nh._nhood_enrichment_helper(
        list(range(20)), f3, ind, ptr, cl.astype(np.uint32), pd.Series(libs), 3, spawn_generators(42, 20)
    ).astype(np.uint32)

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        occ, interval = pp.co_occurrence(mk(), "cluster", copy=True)
    out["cooc_occ"], out["cooc_interval"] = occ, interval
    sp32 = xy.astype(np.float32)
    out["cooc_counts"] = pp._occur_count(sp32[:, 0].copy(), sp32[:, 1].copy(), interval[1:] ** 2, cl.astype(np.int32), 200, 3, 49)
    occ2, interval2 = pp.co_occurrence(mk(), "cluster", interval=np.array([300.0, 10.0, 50.0, 120.5]), copy=True)
    out["cooc_occ_explicit"], out["cooc_interval_explicit"] = occ2, interval2

    for mode in ("L", "F", "G"):
        res = rp.ripley(mk(), "cluster", mode=mode, n_simulations=20, n_observations=300, n_steps=50, seed=7, copy=True)
        out[f"ripley_{mode}_stat"] = res[f"{mode}_stat"]["stats"].to_numpy()
        out[f"ripley_{mode}_sims"] = res["sims_stat"]["stats"].to_numpy()
        out[f"ripley_{mode}_bins"] = res["bins"]
        out[f"ripley_{mode}_pvalues"] = res["pvalues"]
    from scipy.spatial import ConvexHull
    from sklearn.neighbors import KDTree

    hull = ConvexHull(xy.astype(np.float64))
    support = np.linspace(0, (hull.volume / 2) ** 0.5, 50)
    out["ripley_area"] = np.float64(hull.volume)
    out["ripley_support"] = support
    out["ripley_tp"] = np.stack([KDTree(xy[cl == c].astype(np.float64)).two_point_correlation(xy[cl == c].astype(np.float64), support, dualtree=True) for c in range(3)])

    for mode in ("moran", "geary"):
        df = pp.spatial_autocorr(mk(), mode=mode, n_perms=10, seed=3, copy=True, show_progress_bar=False)
        out[f"autocorr_{mode}_index"] = np.array(df.index.tolist())
        out[f"autocorr_{mode}_columns"] = np.array(df.columns.tolist())
        out[f"autocorr_{mode}_values"] = df.to_numpy(dtype=np.float64)
        df = pp.spatial_autocorr(mk(), mode=mode, copy=True, transformation=False, two_tailed=True, genes=["g3", "g1", "g7"])
        out[f"autocorr_{mode}_nt_index"] = np.array(df.index.tolist())
        out[f"autocorr_{mode}_nt_values"] = df.to_numpy(dtype=np.float64)
    df = pp.spatial_autocorr(mk(), mode="moran", attr="obs", genes=["cont"], copy=True)
    out["autocorr_obs_values"] = df.to_numpy(dtype=np.float64)
    from sklearn.preprocessing import normalize

    g = adj.copy()
    normalize(g, norm="l1", axis=1, copy=False)
    out["moments"] = np.array(pp._g_moments(g), dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "dummy_adata.npz"), **out)

    # ------------------------------------------------------------------ config 1: 5 041-spot Visium grid, C=10, n_perms=100
    co = synth.hex_coords(71, 71)
    adj1, _ = nb.GridBuilder(n_neighs=6).build(co)
    lab = synth.categorical_labels(5041, 10, seed=0)
    obs = pd.DataFrame({"cluster": lab.values}, index=[str(i) for i in range(5041)])
    ad = AnnDataLite(obs=obs, obsm={"spatial": co}, obsp={"spatial_connectivities": adj1}, shape=(5041, 0))
    res = nh.nhood_enrichment(ad, "cluster", n_perms=100, seed=42, copy=True, show_progress_bar=False)
    occ, interval = pp.co_occurrence(ad, "cluster", interval=12, copy=True)
    np.savez_compressed(
        os.path.join(OUT, "cfg1_visium5k.npz"), meta=meta, codes=lab.cat.codes.to_numpy().astype(np.int16), nnz=np.int64(adj1.nnz),
        indices_crc=np.int64(__import__("zlib").crc32(adj1.indices.astype("<i4").tobytes())), nhood_z=res.zscore, nhood_count=res.counts,
        cooc_occ=occ, cooc_interval=interval,
    )
    # ------------------------------------------------------------------ jittered co-occurrence / ripley counts (FMA & tie semantics)
    rr = np.random.default_rng(1)
    pts = (rr.random((3000, 2)) * 1000).astype(np.float32)
    lb = rr.integers(0, 4, 3000).astype(np.int32)
    iv = np.linspace(5, 700, 50, dtype=np.float32)
    cref = pp._occur_count(pts[:, 0].copy(), pts[:, 1].copy(), iv[1:] ** 2, lb, 3000, 4, 49)
    P = rr.random((4000, 2)) * 5000
    sup = np.linspace(0, 2500, 50)
    tp = KDTree(P).two_point_correlation(P, sup, dualtree=True)
    gx, gy = np.meshgrid(np.arange(60), np.arange(60))
    Li = np.stack([gx.ravel(), gy.ravel()], 1).astype(np.float64)
    supi = np.linspace(0, 30, 31)
    tpi = KDTree(Li).two_point_correlation(Li, supi, dualtree=True)
    np.savez_compressed(os.path.join(OUT, "pairs_jitter.npz"), meta=meta, pts=pts, labs=lb, interval=iv, cooc_counts=cref, P=P, support=sup, tp=tp,
                        lattice_support=supi, lattice_tp=tpi)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
