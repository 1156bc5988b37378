"""CPU-side tests: key contract, validators, RNG table, sharding, library loading/exports, no-oracle-in-product."""

from __future__ import annotations

import ctypes
import os
import re

import numpy as np
import pandas as pd
import pytest
import scipy.sparse as sp

import squidpy_b200 as sq
from squidpy_b200 import _lib
from squidpy_b200._constants import Key, RipleyStat, SpatialAutocorr
from squidpy_b200._dist import shard_range
from squidpy_b200._rng import spawn_states
from squidpy_b200.gr import _ppatterns as pp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_key_contract():
    # reference tests/graph/test_utils.py:37-68
    assert Key.obsp.spatial_conn() == "spatial_connectivities"
    assert Key.obsp.spatial_conn("foo") == "foo_connectivities"
    assert Key.obsp.spatial_conn("foo_connectivities") == "foo_connectivities"
    assert Key.obsp.spatial_dist() == "spatial_distances"
    assert Key.uns.nhood_enrichment("leiden") == "leiden_nhood_enrichment"
    assert Key.uns.co_occurrence("leiden") == "leiden_co_occurrence"
    assert Key.uns.ripley("leiden", RipleyStat("L")) == "leiden_ripley_L"
    assert str(SpatialAutocorr("moran")) == "moran"
    with pytest.raises(ValueError, match="Invalid option `foo` for `SpatialAutocorr`"):
        SpatialAutocorr("foo")


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "squidpy_b200.h")).read()
    declared = set(re.findall(r"\b(sqb_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"sqb_ctx", "sqb_nhood", "sqb_autocorr", "sqb_status"}
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().sqb_abi_version() == _lib.ABI_VERSION


def test_product_never_touches_oracle():
    pkg = os.path.join(ROOT, "squidpy_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, re.M), f
                assert "liboracle" not in txt, f


def test_no_gpu_fails_loudly():
    try:
        import torch

        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    with pytest.raises(sq.SquidpyB200Error):
        sq.Context(0)


def test_spawn_states_matches_numpy_generators():
    st = spawn_states(42, 5)
    for i, ss in enumerate(np.random.SeedSequence(42).spawn(5)):
        s = np.random.default_rng(ss).bit_generator.state
        assert (int(st[i, 0]) << 64 | int(st[i, 1])) == s["state"]["state"]
        assert (int(st[i, 2]) << 64 | int(st[i, 3])) == s["state"]["inc"]
        assert st[i, 4] == 0
    np.testing.assert_array_equal(spawn_states(42, 5, 2, 4), st[2:4])


def test_shard_range_is_the_reference_chunk_rule():
    # contiguous chunks of ceil(n / n_split): src/squidpy/_utils.py:225-231
    for n, ws in [(1000, 8), (10, 4), (3, 8), (0, 2), (1001, 2)]:
        step = -(-n // ws) if n else 0
        parts = [shard_range(n, r, ws) for r in range(ws)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        for r, (lo, hi) in enumerate(parts):
            assert hi - lo <= step and lo == min(r * step, n)
        assert sum(hi - lo for lo, hi in parts) == n


def test_validators_messages(dummy_adata):
    with pytest.raises(KeyError, match="Cluster key `foo` not found"):
        sq.gr.nhood_enrichment(dummy_adata, "foo")
    with pytest.raises(TypeError, match="to be `categorical`"):
        sq.gr.nhood_enrichment(dummy_adata, "cont")
    with pytest.raises(KeyError, match="Spatial connectivity key `bar_connectivities` not found"):
        sq.gr.nhood_enrichment(dummy_adata, "cluster", connectivity_key="bar")
    with pytest.raises(ValueError, match="Expected `n_perms` to be positive, found `0`"):
        sq.gr.nhood_enrichment(dummy_adata, "cluster", n_perms=0)
    with pytest.raises(KeyError, match="Spatial basis `nope` not found"):
        sq.gr.co_occurrence(dummy_adata, "cluster", spatial_key="nope")
    with pytest.raises(ValueError, match="Expected interval to be of length"):
        sq.gr.co_occurrence(dummy_adata, "cluster", interval=np.array([3.0]))
    with pytest.raises(ValueError, match="Unsupported metric"):
        sq.gr.ripley(dummy_adata, "cluster", mode="L", metric="cosine")
    with pytest.raises(ValueError, match="Invalid option `Z` for `RipleyStat`"):
        sq.gr.ripley(dummy_adata, "cluster", mode="Z")
    with pytest.raises(NotImplementedError, match="adata.foo"):
        sq.gr.spatial_autocorr(dummy_adata, attr="foo")
    with pytest.warns(FutureWarning, match="`n_jobs` of `co_occurrence\\(\\)` is deprecated"):
        with pytest.raises(ValueError):
            sq.gr.co_occurrence(dummy_adata, "cluster", interval=np.array([3.0]), n_jobs=2)


def test_moments_and_pvalues_match_reference_golden(golden_dummy):
    from sklearn.preprocessing import normalize

    g = golden_dummy
    adj = sp.csr_matrix((g["adj_data"], g["adj_indices"], g["adj_indptr"]), shape=(200, 200))
    w = adj.copy()
    normalize(w, norm="l1", axis=1, copy=False)
    s0, s1, s2 = pp._g_moments(w)
    assert isinstance(s0, np.float32)  # float32 moments are part of the reference's p-values
    np.testing.assert_array_equal(np.array([s0, s1, s2], np.float64), g["moments"])
    # whole p-value table of the golden reference run, fed with the golden scores
    cols = list(g["autocorr_moran_columns"])
    tab = g["autocorr_moran_values"]
    params = {"mode": "moran", "two_tailed": False, "expected": -1.0 / 199}
    p, v = pp._analytic_pval(tab[:, cols.index("I")], w, params)
    np.testing.assert_allclose(p, tab[:, cols.index("pval_norm")], rtol=1e-12)
    np.testing.assert_allclose(v, tab[0, cols.index("var_norm")], rtol=1e-7)
    np.testing.assert_allclose(pp._multipletests(tab[:, cols.index("pval_norm")], "fdr_bh"), tab[:, cols.index("pval_norm_fdr_bh")], rtol=1e-12)


def test_var_norm_closed_form(golden_dummy):
    # reference tests/graph/test_ppatterns.py:108-137
    g = golden_dummy
    w = sp.csr_matrix((g["adj_data"], g["adj_indices"], g["adj_indptr"]), shape=(200, 200))
    s0, s1, s2 = (float(v) for v in pp._g_moments(w))
    n = 200
    _, v_m = pp._analytic_pval(np.zeros(3), w, {"mode": "moran", "two_tailed": False, "expected": -1 / 199})
    exp_m = (n * n * s1 - n * s2 + 3 * s0 * s0) / ((n - 1) * (n + 1) * s0 * s0) - (1.0 / (n - 1)) ** 2
    np.testing.assert_allclose(v_m, exp_m, rtol=1e-6)
    _, v_g = pp._analytic_pval(np.zeros(3), w, {"mode": "geary", "two_tailed": False, "expected": 1.0})
    exp_g = ((2 * s1 + s2) * (n - 1) - 4 * s0 * s0) / (2 * (n + 1) * s0 * s0)
    np.testing.assert_allclose(v_g, exp_g, rtol=1e-6)


def test_multipletests_variants():
    p = np.array([0.01, 0.04, 0.03, 0.5, 0.2])
    np.testing.assert_allclose(pp._multipletests(p, "bonferroni"), np.minimum(p * 5, 1))
    bh = pp._multipletests(p, "fdr_bh")
    assert (bh >= p).all() and (bh <= 1).all() and bh[np.argmin(p)] == pytest.approx(0.05)
    holm = pp._multipletests(p, "holm")
    assert holm[0] == pytest.approx(0.05) and (holm <= 1).all()


def test_anndata_lite_indexing():
    X = sp.random(6, 4, density=0.5, format="csr", random_state=0)
    ad = sq.AnnDataLite(X=X, var=pd.DataFrame({"highly_variable": [True, False, True, False]}, index=list("abcd")))
    sub = ad[:, ["c", "a"]]
    assert sub.shape == (6, 2) and list(sub.var_names) == ["c", "a"]
    np.testing.assert_array_equal(sub.X.toarray(), X.toarray()[:, [2, 0]])
    assert list(ad[:, ad.var["highly_variable"]].var_names) == ["a", "c"]
    with pytest.raises(KeyError):
        ad[:, ["zz"]]


@pytest.mark.parametrize("seed", [0, 7, 2**32 - 1, 2**32, 2**70 + 12345, 123456789012345678901234567890, [1, 2, 3], [2**40, 5]])
def test_spawn_states_vectorised_equals_numpy(seed):
    """The array implementation of SeedSequence.spawn + PCG64 seeding (squidpy_b200/_rng.py) against numpy objects."""
    from squidpy_b200._rng import generator_state, spawn_generators, spawn_states

    n = 67
    exp = np.array([generator_state(g) for g in spawn_generators(seed, n)], dtype=np.uint64)
    np.testing.assert_array_equal(spawn_states(seed, n), exp)
    np.testing.assert_array_equal(spawn_states(seed, n, 13, 40), exp[13:40])


def test_spawn_states_fresh_entropy():
    from squidpy_b200._rng import spawn_states

    a, b = spawn_states(None, 4), spawn_states(None, 4)
    assert a.shape == (4, 6) and not np.array_equal(a, b)  # seed=None draws OS entropy, like the reference


def test_ppp_block_sampling_is_the_scalar_stream():
    """``_ppp`` draws candidates in blocks; points AND the generator position afterwards must equal the reference's
    one-candidate-at-a-time loop (``_ripley.py:255-269``)."""
    from scipy.spatial import ConvexHull, Delaunay

    from squidpy_b200.gr._ripley import _ppp

    def scalar(hull, n_sim, n_obs, rng):
        vxs = hull.points[hull.vertices]
        deln = Delaunay(vxs)
        bbox = np.array([*vxs.min(0), *vxs.max(0)])
        out = np.empty((n_sim, n_obs, 2))
        for i in range(n_sim):
            k = 0
            while k < n_obs:
                x, y = rng.uniform(bbox[0], bbox[2]), rng.uniform(bbox[1], bbox[3])
                if deln.find_simplex((x, y)) >= 0:
                    out[i, k] = (x, y)
                    k += 1
        return out.squeeze()

    pts = np.random.default_rng(0).normal(size=(400, 2)) * [3, 1] + [10, -4]
    hull = ConvexHull(pts)
    for n_obs in (1, 7, 100, 700):
        a, b = np.random.default_rng(5), np.random.default_rng(5)
        np.testing.assert_array_equal(scalar(hull, 2, n_obs, a), _ppp(hull, 2, n_obs, b))
        assert a.random() == b.random()


def test_gathered_blocks_are_in_global_permutation_order():
    """`_dist.gathered_stats_device` relies on this: every rank pads its block of rows to `step = ceil(n / world)` rows, the blocks
    are concatenated in rank order (all_gather_into_tensor), and the FIRST n rows of the result are then the rows of the global
    job in order -- because `shard_range` hands out contiguous blocks of `step` rows and only the last non-empty one is short."""
    from squidpy_b200._dist import shard_range

    for n in (1, 2, 7, 8, 9, 100, 1000, 1001):
        for ws in (1, 2, 3, 4, 8, 16):
            step = -(-n // ws)
            full = np.full((ws * step,), -1, dtype=np.int64)
            covered = 0
            for r in range(ws):
                lo, hi = shard_range(n, r, ws)
                assert 0 <= lo <= hi <= n and hi - lo <= step
                assert lo == min(r * step, n)  # contiguous blocks of `step`
                full[r * step : r * step + (hi - lo)] = np.arange(lo, hi)
                covered += hi - lo
            assert covered == n
            np.testing.assert_array_equal(full[:n], np.arange(n))
            assert (full[n:] == -1).all()


def test_row_record_count_identity():
    """The algorithm of nhood_count_recs_kernel, restated in numpy and held against the oracle: the stored entries with j >= i of a
    structurally symmetric graph are cut into row records {i, j0, j1, j2} (unused slots point at the spare node n, which carries
    the spare label C); ONE increment U[l_i][l_j] per slot into a histogram with C + 1 columns; the flush adds the mirror,
    counts = U + U^T over the first C columns, and every self loop is taken off the diagonal once.  Directed graphs: records of
    all entries, no mirror."""
    import scipy.sparse as sp

    from oracle import ref

    rng = np.random.default_rng(12)
    for trial in range(20):
        n, C = int(rng.integers(5, 300)), int(rng.integers(2, 9))
        a = sp.random(n, n, density=float(rng.uniform(0.01, 0.2)), format="csr", random_state=int(rng.integers(1 << 30)), dtype=np.float32)
        sym = trial % 2 == 0
        if sym:
            a = ((a + a.T) > 0).astype(np.float32) + sp.diags((rng.random(n) < 0.3).astype(np.float32))
        a = sp.csr_matrix(a)
        a.eliminate_zeros()
        a.sort_indices()
        lab = rng.integers(0, C, n)
        lab1 = np.append(lab, C)  # the spare row
        recs = []
        for i in range(n):
            cols = a.indices[a.indptr[i] : a.indptr[i + 1]]
            if sym:
                cols = cols[cols >= i]
            for k in range(0, len(cols), 3):
                slot = list(cols[k : k + 3]) + [n] * (3 - len(cols[k : k + 3]))
                recs.append((i, *slot))
        hist = np.zeros((C, C + 1), dtype=np.int64)
        for i, j0, j1, j2 in recs:
            for j in (j0, j1, j2):
                hist[lab1[i], lab1[j]] += 1
        u = hist[:, :C]
        if sym:
            got = u + u.T
            for i in np.flatnonzero(a.diagonal() != 0):
                got[lab[i], lab[i]] -= 1
        else:
            got = u
        exp = ref.nhood_count(a.indptr, a.indices, lab.astype(np.uint32), C)
        np.testing.assert_array_equal(got, exp)
