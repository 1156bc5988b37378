"""The CPU oracle against the committed golden vectors (generated from the running reference, see
tests/golden/make_golden.py) and against the known-answer vectors of SURVEY.md Appendix A."""

from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from oracle import ref


def _graph(g):
    return g["adj_indptr"].astype(np.uint32), g["adj_indices"].astype(np.uint32)


def test_kat1_count_kernel():
    # reference tests/conftest.py:177-194 + tests/graph/test_nhood.py:158 (unweighted expectation)
    out = ref.nhood_count([0, 2, 3, 5, 7, 9], [1, 2, 4, 0, 1, 1, 4, 2, 3], [0, 0, 0, 1, 1], 2)
    assert out.dtype == np.uint32
    np.testing.assert_array_equal(out, [[4, 1], [2, 2]])


def test_kat2_nhood(golden_dummy):
    g = golden_dummy
    ptr, ind = _graph(g)
    z, count, perms = ref.nhood_enrichment(ptr, ind, g["cl"].astype(np.uint32), 3, 42, 20)
    np.testing.assert_array_equal(count, [[184, 128, 150], [140, 111, 115], [143, 105, 124]])
    np.testing.assert_array_equal(perms[0], [[197, 121, 144], [127, 122, 117], [142, 114, 116]])
    np.testing.assert_array_equal(perms[19], [[189, 119, 154], [129, 118, 119], [151, 107, 114]])
    np.testing.assert_array_equal(perms, g["nhood_perms"])
    np.testing.assert_array_equal(z, g["nhood_z"])
    assert z[0, 0] == 0.4957188836779417


def test_kat2_nhood_library(golden_dummy):
    g = golden_dummy
    ptr, ind = _graph(g)
    lib = np.repeat([0, 1], 100)
    z, count, perms = ref.nhood_enrichment(ptr, ind, g["cl"].astype(np.uint32), 3, 42, 20, lib_codes=lib, n_libs=2)
    np.testing.assert_array_equal(perms[0], [[167, 152, 143], [154, 99, 113], [147, 120, 105]])
    np.testing.assert_array_equal(perms, g["nhood_lib_perms"])
    np.testing.assert_array_equal(z, g["nhood_lib_z"])


def test_kat3_cooccurrence(golden_dummy):
    g = golden_dummy
    sp32 = g["xy"].astype(np.float32)
    interval = g["cooc_interval"]
    assert interval.dtype == np.float32 and interval.size == 50
    occ, counts = ref.co_occurrence_helper(sp32[:, 0], sp32[:, 1], interval, g["cl"].astype(np.int32))
    np.testing.assert_array_equal(counts, g["cooc_counts"])
    np.testing.assert_array_equal(counts[:, :, 0], [[138, 119, 110], [119, 114, 92], [110, 92, 98]])
    assert counts.sum() == 572934
    np.testing.assert_allclose(occ, g["cooc_occ"], rtol=1e-12)


def test_cooc_fma_semantics(golden_pairs):
    g = golden_pairs
    thr = g["interval"][1:] ** 2
    c_fma = ref.occur_count(g["pts"][:, 0], g["pts"][:, 1], thr, g["labs"], 4, use_fma=True)
    np.testing.assert_array_equal(c_fma, g["cooc_counts"])  # the reference JIT contracts dx*dx + dy*dy into an FMA
    c_full = ref.occur_count(g["pts"][:, 0], g["pts"][:, 1], thr, g["labs"], 4, use_fma=True, compact=False)
    np.testing.assert_array_equal(c_full, c_fma)


def test_kat4_ripley(golden_dummy, golden_pairs):
    g = golden_dummy
    for c in range(3):
        pts = g["xy"][g["cl"] == c].astype(np.float64)
        np.testing.assert_array_equal(ref.pair_counts(pts, g["ripley_support"]), g["ripley_tp"][c])
    np.testing.assert_array_equal(g["ripley_tp"][0][:6], [77, 79, 83, 93, 123, 157])
    _, L = ref.l_function(g["xy"][g["cl"] == 0].astype(np.float64), g["ripley_support"], 200, float(g["ripley_area"]))
    np.testing.assert_allclose(L[:5], [0, 1.9322422482087205, 3.346741746428617, 5.465206386414105, 9.266708284636021], rtol=1e-14)
    p = golden_pairs
    np.testing.assert_array_equal(ref.pair_counts(p["P"], p["support"]), p["tp"])
    gx, gy = np.meshgrid(np.arange(60), np.arange(60))
    lattice = np.stack([gx.ravel(), gy.ravel()], 1).astype(np.float64)
    np.testing.assert_array_equal(ref.pair_counts(lattice, p["lattice_support"]), p["lattice_tp"])  # exact ties


def test_moran_geary_restatement(golden_dummy):
    """scanpy is absent: the restatement is only cross-checked against an independent scipy formulation
    (parity unpinned)."""
    from sklearn.preprocessing import normalize

    g = golden_dummy
    adj = sp.csr_matrix((g["adj_data"], g["adj_indices"], g["adj_indptr"]), shape=(200, 200))
    w = adj.copy()
    normalize(w, norm="l1", axis=1, copy=False)
    vals = g["X"].T
    I = ref.morans_i(w, vals)
    chk = np.array([ref.morans_i_dense_check(w, v) for v in vals])
    np.testing.assert_allclose(I, chk, rtol=1e-10, atol=1e-13)
    Xs = sp.random(200, 40, density=0.1, random_state=1, format="csr", dtype=np.float32)
    np.testing.assert_allclose(ref.morans_i(w, Xs.T.tocsr()), ref.morans_i(w, Xs.T.toarray()), rtol=1e-13)
    wd = w.astype(np.float64).tocoo()
    C = ref.gearys_c(w, vals)
    chk = [(199) * np.sum(wd.data * (v[wd.row] - v[wd.col]) ** 2) / (2 * wd.data.sum() * np.sum((v - v.mean()) ** 2)) for v in vals]
    np.testing.assert_allclose(C, chk, rtol=1e-10)
    assert np.isnan(ref.morans_i(w, np.ones((1, 200)))[0])
    # row-permuted W (g[idx, :]) against scipy row indexing
    idx = np.random.default_rng(0).permutation(200)
    np.testing.assert_allclose(ref.morans_i(w, vals, row_perm=idx), ref.morans_i(w[idx, :], vals), rtol=1e-12)
    # the golden autocorr table (reference driver + this restatement) keeps its first column equal to the restatement
    cols = list(g["autocorr_moran_columns"])
    tab = g["autocorr_moran_values"]
    order = [int(s[1:]) for s in g["autocorr_moran_index"]]
    np.testing.assert_allclose(tab[:, cols.index("I")], I[order], rtol=1e-12)


def test_rng_replay_vs_numpy():
    for seed, n in [(0, 2), (42, 7), (12345, 1000), (3, 5041)]:
        st = ref.spawn_states(seed, 3)
        gens = [np.random.default_rng(s) for s in np.random.SeedSequence(seed).spawn(3)]
        for k in range(3):
            a = np.arange(n, dtype=np.uint32)
            exp = a.copy()
            gens[k].shuffle(exp)
            s6 = st[k].copy()
            np.testing.assert_array_equal(ref.shuffle_u32(s6, a), exp)
            # the stream continues identically (consecutive shuffles from one generator, _shuffle_group)
            exp2 = a.copy()
            gens[k].shuffle(exp2)
            np.testing.assert_array_equal(ref.shuffle_u32(s6, a), exp2)
    st = ref.spawn_states(5, 1)[0]
    gen = np.random.default_rng(np.random.SeedSequence(5).spawn(1)[0])
    np.testing.assert_array_equal(ref.permutation(st, 1234), gen.permutation(1234))


def test_ligrec_restatement_matches_reference_golden():
    """oracle.ref.ligrec_counts (numpy) against p-values produced by the reference's own _analysis (tests/golden/ligrec.npz)."""
    import os

    from tests.golden.make_golden_ligrec import CASES, frame, make_case

    gold = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "ligrec.npz"), allow_pickle=False))
    for name, (seed, n_cells, n_genes, n_cls, kind, thr, n_perms, pseed) in CASES.items():
        x, cl, inter, cpairs = make_case(seed, n_cells, n_genes, n_cls, kind)
        g = frame(x, cl, n_cls).groupby("clusters", observed=True)
        mean_obs = g.mean().values
        inv = 1.0 / np.maximum(g.size().values.astype(np.float64), 1)
        valid = ~np.isnan(gold[f"{name}_pvalues"])
        counts = ref.ligrec_counts(x, cl, n_cls, ref.spawn_states(pseed, n_perms), inv, mean_obs, inter, cpairs, valid)
        np.testing.assert_array_equal(counts[valid] / n_perms, gold[f"{name}_pvalues"][valid])


def test_sepal_restatement_matches_reference_golden():
    """oracle.ref.sepal_score (numpy, no fastmath) against scores of the reference's numba kernel (tests/golden/sepal.npz)."""
    import os

    from squidpy_b200.gr._sepal import _compute_idxs
    from tests.golden.make_golden_sepal import make_case

    gold = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "sepal.npz"), allow_pickle=False))
    g, co, k, vals = make_case("square")
    sat, si, un, ui = _compute_idxs(g, co, k)
    np.testing.assert_array_equal(si, gold["square_sat_idx"])
    np.testing.assert_array_equal(ui, gold["square_unsat_idx"])
    for q in (10, 11, 13, 14, 15):  # the quickly converging genes (the loop is interpreted)
        got = ref.sepal_score(vals[:, q], False, 30000, sat, si, un, ui)
        assert abs(got - gold["square_score"][q]) <= 2.5e-3, (q, got, gold["square_score"][q])
