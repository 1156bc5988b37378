"""numpy-level CPU oracle (TEST INFRASTRUCTURE ONLY — see oracle/__init__.py).

Thin ctypes wrappers around ``liboracle.so`` plus numpy restatements of the host-side float post-processing
of every reference function on the hot path.  Each function cites the reference ``file:line`` (relative to
``/root/reference/``) it follows.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib as _lib

_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


# ---------------------------------------------------------------------------------------------------
# RNG: src/squidpy/_utils.py:240-241  spawn_generators(seed, n)
# ---------------------------------------------------------------------------------------------------
def spawn_states(seed, n: int) -> np.ndarray:
    """``[default_rng(s) for s in SeedSequence(seed).spawn(n)]`` flattened to an (n, 6) uint64 table:
    state_hi, state_lo, inc_hi, inc_lo, has_uint32, uinteger."""
    out = np.empty((n, 6), dtype=np.uint64)
    m64 = (1 << 64) - 1
    for i, ss in enumerate(np.random.SeedSequence(seed).spawn(n)):
        st = np.random.default_rng(ss).bit_generator.state
        s, inc = st["state"]["state"], st["state"]["inc"]
        out[i] = (s >> 64, s & m64, inc >> 64, inc & m64, st["has_uint32"], st["uinteger"])
    return out


def shuffle_u32(state6: np.ndarray, arr: np.ndarray) -> np.ndarray:
    """numpy ``Generator.shuffle`` replay on a uint32 vector; updates ``state6`` in place."""
    a = _c(arr, np.uint32).copy()
    f = _lib().orc_shuffle_u32
    f.argtypes = [_u64p, _u32p, C.c_int64]
    f.restype = None
    f(state6, a, a.size)
    return a


def permutation(state6: np.ndarray, n: int) -> np.ndarray:
    a = np.empty(n, dtype=np.int64)
    f = _lib().orc_permutation_i64
    f.argtypes = [_u64p, _i64p, C.c_int64]
    f.restype = None
    f(state6, a, n)
    return a


def pcg64_raw(state6: np.ndarray, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.uint64)
    f = _lib().orc_pcg64_raw
    f.argtypes = [_u64p, _u64p, C.c_int64]
    f.restype = None
    f(state6, out, n)
    return out


def _groups(lib_codes, n_libs):
    """member indices per library category, category order (gr/_utils.py:208-209 np.where(libraries == c))."""
    if lib_codes is None:
        return np.zeros(1, np.int64), np.zeros(2, np.int64), 0
    lib_codes = np.asarray(lib_codes)
    idx = [np.where(lib_codes == c)[0].astype(np.int64) for c in range(n_libs)]
    ptr = np.zeros(n_libs + 1, np.int64)
    ptr[1:] = np.cumsum([len(i) for i in idx])
    return (np.concatenate(idx) if ptr[-1] else np.zeros(1, np.int64)), ptr, n_libs


# ---------------------------------------------------------------------------------------------------
# nhood_enrichment: src/squidpy/gr/_nhood.py
# ---------------------------------------------------------------------------------------------------
def nhood_count(indptr, indices, labels, n_cls: int) -> np.ndarray:
    """_nenrich_{n_cls} (:54-141)."""
    indptr, indices, labels = _c(indptr, np.uint32), _c(indices, np.uint32), _c(labels, np.uint32)
    n = indptr.size - 1
    out = np.empty((n_cls, n_cls), np.uint32)
    scratch = np.empty(max(n * n_cls, 1), np.uint32)
    f = _lib().orc_nhood_count
    f.argtypes = [_u32p, _u32p, _u32p, C.c_int64, C.c_int, _u32p, _u32p]
    f.restype = None
    f(indptr, indices, labels, n, n_cls, scratch, out)
    return out


def nhood_perm_counts(indptr, indices, base, n_cls, states, lib_codes=None, n_libs=0, n_threads=0) -> np.ndarray:
    """_nhood_enrichment_helper (:516-547) for all permutations -> uint32 (P, C, C)."""
    indptr, indices, base = _c(indptr, np.uint32), _c(indices, np.uint32), _c(base, np.uint32)
    states = _c(states, np.uint64)
    n, p = indptr.size - 1, states.shape[0]
    gi, gp, ng = _groups(lib_codes, n_libs)
    out = np.empty((p, n_cls, n_cls), np.uint32)
    f = _lib().orc_nhood_perms
    f.argtypes = [_u32p, _u32p, _u32p, C.c_int64, C.c_int, _u64p, C.c_int64, _i64p, _i64p, C.c_int, _u32p, C.c_int]
    f.restype = None
    f(indptr, indices, base, n, n_cls, states, p, gi, gp, ng, out, n_threads)
    return out


def shuffle_labels(base, states, lib_codes=None, n_libs=0) -> np.ndarray:
    base, states = _c(base, np.uint32), _c(states, np.uint64)
    gi, gp, ng = _groups(lib_codes, n_libs)
    out = np.empty((states.shape[0], base.size), np.uint32)
    f = _lib().orc_shuffle_labels
    f.argtypes = [_u32p, C.c_int64, _u64p, C.c_int64, _i64p, _i64p, C.c_int, _u32p]
    f.restype = None
    f(base, base.size, states, states.shape[0], gi, gp, ng, out)
    return out


def nhood_enrichment(indptr, indices, labels, n_cls, seed, n_perms, lib_codes=None, n_libs=0, n_threads=0):
    """nhood_enrichment core (:205-231): returns (zscore f64[C,C], count u32[C,C], perms u32[P,C,C])."""
    count = nhood_count(indptr, indices, labels, n_cls)
    states = spawn_states(seed, n_perms)
    perms = nhood_perm_counts(indptr, indices, labels, n_cls, states, lib_codes, n_libs, n_threads)
    pf = perms.astype(np.float64)  # perms array is float64 in the reference (:526)
    with np.errstate(divide="ignore", invalid="ignore"):
        zscore = (count - pf.mean(axis=0)) / pf.std(axis=0)  # (:231), no zero-std guard
    return zscore, count, perms


# ---------------------------------------------------------------------------------------------------
# co_occurrence: src/squidpy/gr/_ppatterns.py:283-358
# ---------------------------------------------------------------------------------------------------
def occur_count(x, y, thr, labs, k, use_fma=True, compact=True, n_threads=0) -> np.ndarray:
    """_occur_count (:283-310) -> int64 (k, k, L)."""
    x, y, thr, labs = _c(x, np.float32), _c(y, np.float32), _c(thr, np.float32), _c(labs, np.int32)
    L = thr.size
    out = np.empty((k, k, L), np.int64)
    f = _lib().orc_cooc_counts
    f.argtypes = [_f32p, _f32p, C.c_int64, _i32p, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, _i64p, C.c_int]
    f.restype = C.c_int
    if f(x, y, x.size, labs, k, thr, L, int(use_fma), int(compact), out, n_threads) != 0:
        raise MemoryError("oracle co-occurrence scratch")
    return out


def co_occurrence_helper(x, y, interval, labs, use_fma=True, compact=True, n_threads=0):
    """_co_occurrence_helper (:313-358) -> (occ f64[k,k,L], counts i64[k,k,L])."""
    interval = _c(interval, np.float32)
    labs = _c(labs, np.int32)
    k = len(np.unique(labs))
    L = interval.size - 1
    thr = interval[1:] ** 2  # float32
    counts = occur_count(x, y, thr, labs, k, use_fma, compact, n_threads)
    occ = np.zeros((k, k, L), np.float64)
    row_sums = counts.sum(axis=0)  # (k, L) int64
    totals = row_sums.sum(axis=0)  # (L,)
    for r in range(L):
        with np.errstate(divide="ignore", invalid="ignore"):
            probs = row_sums[:, r] / totals[r]
        for c in range(k):
            for i in range(k):
                if probs[i] != 0.0 and row_sums[c, r] != 0.0:
                    occ[i, c, r] = (counts[c, i, r] / row_sums[c, r]) / probs[i]
    return occ, counts


# ---------------------------------------------------------------------------------------------------
# ripley: src/squidpy/gr/_ripley.py:212-227
# ---------------------------------------------------------------------------------------------------
def pair_counts(points, support, n_threads=0) -> np.ndarray:
    """KDTree.two_point_correlation(points, support) restated by brute force (self pairs included)."""
    pts, r = _c(points, np.float64), _c(support, np.float64)
    out = np.empty(r.size, np.int64)
    f = _lib().orc_pair_counts_f64
    f.argtypes = [_f64p, C.c_int64, _f64p, C.c_int, _i64p, C.c_int]
    f.restype = None
    f(pts, pts.shape[0], r, r.size, out, n_threads)
    return out


def l_function(points, support, n, area, n_threads=0):
    """_l_function (:212-227)."""
    m = np.asarray(points).shape[0]
    cnt = pair_counts(points, support, n_threads) - m
    intensity = n / area
    k_estimate = (cnt / n) / intensity
    return np.asarray(support), np.sqrt(k_estimate / np.pi)


# ---------------------------------------------------------------------------------------------------
# Moran / Geary: scanpy.metrics restatement (parity unpinned) — call sites _ppatterns.py:200,205,216,267,272
# ---------------------------------------------------------------------------------------------------
def _autocorr(mode, g, vals, row_perm=None, n_threads=0):
    import scipy.sparse as sp

    g = sp.csr_matrix(g)
    wp, wi = _c(g.indptr, np.int32), _c(g.indices, np.int32)
    wd = _c(g.data, np.float64)  # scanpy: g.data.astype(float64)
    n = g.shape[0]
    rp = None if row_perm is None else _c(row_perm, np.int64)
    rp_arg = rp.ctypes.data_as(C.c_void_p) if rp is not None else None
    if sp.issparse(vals):
        v = sp.csr_matrix(vals)
        nf = v.shape[0]
        out = np.empty(nf, np.float64)
        f = _lib().orc_autocorr_csr
        f.argtypes = [C.c_int, _i32p, _i32p, _f64p, C.c_int64, _i64p, _i32p, _f64p, C.c_int64, C.c_void_p, _f64p, C.c_int]
        f.restype = None
        f(mode, wp, wi, wd, n, _c(v.indptr, np.int64), _c(v.indices, np.int32), _c(v.data, np.float64), nf, rp_arg, out, n_threads)
        return out
    v = np.atleast_2d(_c(vals, np.float64))
    out = np.empty(v.shape[0], np.float64)
    f = _lib().orc_autocorr_dense
    f.argtypes = [C.c_int, _i32p, _i32p, _f64p, C.c_int64, _f64p, C.c_int64, C.c_void_p, _f64p, C.c_int]
    f.restype = None
    f(mode, wp, wi, wd, n, v, v.shape[0], rp_arg, out, n_threads)
    return out


def morans_i(g, vals, row_perm=None, n_threads=0):
    return _autocorr(0, g, vals, row_perm, n_threads)


def gearys_c(g, vals, row_perm=None, n_threads=0):
    return _autocorr(1, g, vals, row_perm, n_threads)


def morans_i_dense_check(g, x):
    """Independent formulation used to cross-check the restatement: z^T (W z) with scipy.sparse @."""
    x = np.asarray(x, np.float64)
    z = x - x.mean()
    gd = g.astype(np.float64)
    return x.size / gd.sum() * float(z @ (gd @ z)) / float(z @ z)


def multipletests_fdr_bh(pvals):
    """Benjamini-Hochberg as statsmodels.stats.multitest.multipletests(method='fdr_bh') returns it
    (second element), used at _ppatterns.py:242-245."""
    p = np.asarray(pvals, np.float64)
    n = p.size
    order = np.argsort(p)
    ps = p[order]
    ecdf = np.arange(1, n + 1) / float(n)
    corrected = ps / ecdf
    corrected = np.minimum.accumulate(corrected[::-1])[::-1]
    corrected[corrected > 1] = 1
    out = np.empty(n)
    out[order] = corrected
    return out


# ---------------------------------------------------------------------------------------------------
# interaction_matrix: src/squidpy/gr/_nhood.py:349-429  (SURVEY 8f item 2 — oracle prepared ahead of the kernel)
# ---------------------------------------------------------------------------------------------------
def interaction_matrix(graph, codes, n_cats: int, weights: bool = False, normalized: bool = False) -> np.ndarray:
    """``interaction_matrix`` restated with numpy (:383-404 + ``_interaction_matrix`` :412-429).

    ``graph``: scipy CSR (N x N); ``codes``: int category codes with -1 for NaN labels (pandas convention) — observations
    with NaN labels are removed from rows AND columns (:388-395).  output[a, b] = sum over stored entries (i -> j) of the
    restricted graph with code(i)=a, code(j)=b of the entry's value (``weights=True``) or of 1 (``weights=False``);
    dtype int64 for bool/integer graphs else float64 (:398); ``normalized`` divides every row by its sum (:403-404).
    The reference accumulates in CSR order; ``np.add.at`` does the same, so float sums are bit-identical."""
    import pandas as pd
    import scipy.sparse as sp

    g = sp.csr_matrix(graph)
    codes = np.asarray(codes)
    mask = codes >= 0
    if not mask.any():
        raise RuntimeError("After removing NaNs, none remain.")
    if not mask.all():
        g = g[mask, :][:, mask]
        codes = codes[mask]
    dtype = np.int64 if (pd.api.types.is_bool_dtype(g.dtype) or pd.api.types.is_integer_dtype(g.dtype)) else np.float64
    out = np.zeros((n_cats, n_cats), dtype=dtype)
    rows = np.repeat(codes, np.diff(g.indptr))
    cols = codes[g.indices]
    vals = g.data.astype(dtype) if weights else np.ones(g.data.size, dtype=dtype)
    np.add.at(out, (rows, cols), vals)
    if normalized:
        out = out / out.sum(axis=1).reshape((-1, 1))
    return out


# ---------------------------------------------------------------------------------------------------
# ligrec: src/squidpy/gr/_ligrec.py:616-676 (_score_permutations)
# ---------------------------------------------------------------------------------------------------
def ligrec_counts(data, clustering, n_cls, states, inv_counts, mean_obs, interactions, interaction_clusters, valid) -> np.ndarray:
    """``_score_permutations`` restated with numpy: per permutation shuffle the cluster labels (exact numpy stream),
    accumulate ``groups[cl, g] += data[cell, g]`` over the cells in ascending order (``np.add.at`` is unbuffered and
    in-order, so the float64 sums are the reference's), scale by ``inv_counts`` and count ``shuf > obs``."""
    data = np.asarray(data, np.float64)
    labs = shuffle_labels(np.asarray(clustering, np.uint32), states)
    rec, lig = interactions[:, 0], interactions[:, 1]
    a, b = interaction_clusters[:, 0], interaction_clusters[:, 1]
    obs = mean_obs[a][:, rec].T + mean_obs[b][:, lig].T
    counts = np.zeros((interactions.shape[0], interaction_clusters.shape[0]), np.int64)
    for p in range(labs.shape[0]):
        groups = np.zeros((n_cls, data.shape[1]), np.float64)
        np.add.at(groups, labs[p].astype(np.int64), data)
        groups *= np.asarray(inv_counts)[:, None]
        shuf = groups[a][:, rec].T + groups[b][:, lig].T
        counts += ((shuf > obs) & valid).astype(np.int64)
    return counts


# ---------------------------------------------------------------------------------------------------
# sepal: src/squidpy/gr/_sepal.py:236-289 (_diffusion)
# ---------------------------------------------------------------------------------------------------
def sepal_score(conc, use_hex, n_iter, sat, sat_idx, unsat, unsat_idx, dt=0.001, thresh=1e-8) -> float:
    """``_diffusion`` restated with numpy (float64, no fastmath): dt * first iteration whose entropy change is <= thresh."""
    conc = np.array(conc, dtype=np.float64)
    eps = np.finfo(np.float64).eps
    prev = 1.0
    for i in range(n_iter):
        nhood = conc[sat_idx].sum(axis=1)
        c = conc[sat]
        d2 = (2.0 * nhood - 12.0 * c) / 3.0 if use_hex else nhood - 4.0 * c
        dcdt = np.zeros_like(conc)
        dcdt[sat] = d2
        conc[sat] += dcdt[sat] * dt
        conc[unsat] += dcdt[unsat_idx] * dt
        conc[conc < 0] = 0
        x = conc[sat]
        x = x[x > 0]
        s = x.sum()
        ent = 0.0 if s < eps else float((-np.log(np.maximum(x / s, eps)) * (x / s)).sum())
        ent /= sat.shape[0]
        if abs(ent - prev) <= thresh:
            return dt * i
        prev = ent
    return float("nan")
