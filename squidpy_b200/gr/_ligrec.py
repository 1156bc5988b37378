"""``ligrec`` — receptor-ligand permutation test (CellPhoneDB) with the permutations on a B200; drop-in for the
explicit-interactions use of ``squidpy.gr.ligrec`` (``src/squidpy/gr/_ligrec.py``).

What runs where
  * per permutation: shuffle the cluster labels with numpy's generator ``p`` of ``spawn_generators(seed, n_perms)`` (exact
    device replay, shared with nhood_enrichment), re-form the per-cluster mean of every gene, compare
    ``mean[a, g0] + mean[b, g1] > observed`` for every (interaction, cluster pair) — ``sqb_ligrec_counts``, replacing the
    numba kernel ``_score_permutations`` (``_ligrec.py:616-676``) with the same float64 operation order;
  * once per call, on the host with the reference's own pandas expressions (so that the observed means the decisions are
    taken against are bit-identical): group means, expression-fraction mask, validity, result frames, FDR
    (``_analysis`` ``_ligrec.py:679-776``, ``PermutationTest.test`` ``:230-370``).
Interactions must be given (DataFrame / dict / sequences); the reference's ``interactions=None`` downloads them from
omnipath, which needs network access and is not replicated.
"""

from __future__ import annotations

import ctypes as C
import time
from collections.abc import Iterable, Mapping
from itertools import product
from typing import Any, NamedTuple

import numpy as np
import pandas as pd

from .._constants import ComplexPolicy, CorrAxis, Key
from .._dist import shared_seed
from .._lib import Context, check, default_context, load
from .._rng import spawn_states
from .._validators import assert_categorical_obs, assert_positive, extract_adata_if_sdata
from ._utils import _save_data, logg

__all__ = ["ligrec", "ligrec_analysis", "TempResult"]

SOURCE = "source"
TARGET = "target"


class TempResult(NamedTuple):
    """``(n_interactions, n_interaction_clusters)`` means and p-values (``_ligrec.py:48-54``)."""

    means: np.ndarray
    pvalues: np.ndarray


def ligrec_analysis(data: pd.DataFrame, interactions: np.ndarray, interaction_clusters: np.ndarray, threshold: float = 0.1,
                    n_perms: int = 1000, seed: int | None = None, *, ctx: Context | None = None) -> TempResult:
    """``_analysis`` (``_ligrec.py:679-776``): ``data`` has one integer-named column per gene plus a categorical column
    ``'clusters'`` whose categories are ``0 .. n_cls-1``; ``interactions`` holds column positions, ``interaction_clusters``
    cluster codes."""
    clustering = np.array(data["clusters"].values, dtype=np.int32)
    data = data.astype({c: np.float64 for c in data.columns if c != "clusters"})
    groups = data.groupby("clusters", observed=True)
    mean_obs = groups.mean().values  # (n_clusters, n_genes): pandas' compensated group mean, the reference value
    mask = groups.apply(lambda c: ((c > 0).astype(np.int64).sum() / len(c)) >= threshold).values
    cluster_sizes = groups.size().values.astype(np.float64)
    inv_counts = 1.0 / np.maximum(cluster_sizes, 1)
    data_arr = np.array(data[data.columns.difference(["clusters"])].values, dtype=np.float64, order="C")

    inter = np.ascontiguousarray(interactions, dtype=np.int32)
    cpairs = np.ascontiguousarray(interaction_clusters, dtype=np.int32)
    rec, lig, c1, c2 = inter[:, 0], inter[:, 1], cpairs[:, 0], cpairs[:, 1]
    m_rec = mean_obs[c1, :][:, rec].T
    m_lig = mean_obs[c2, :][:, lig].T
    nonzero = (m_rec > 0) & (m_lig > 0)
    valid = nonzero & mask[c1, :][:, rec].T & mask[c2, :][:, lig].T
    res_means = np.where(nonzero, (m_rec + m_lig) / 2.0, 0.0)

    n_cells, n_genes = data_arr.shape
    n_cls = mean_obs.shape[0]
    lib = load()
    ctx = ctx or default_context()
    counts = np.zeros((inter.shape[0], cpairs.shape[0]), dtype=np.int64)
    valid_u8 = np.ascontiguousarray(valid, dtype=np.uint8)
    mean_c = np.ascontiguousarray(mean_obs, dtype=np.float64)
    with ctx.lock:
        h = C.c_void_p()
        empty_ptr = np.zeros(n_cells + 1, dtype=np.uint32)
        # a graph-less handle: only the label shuffle of the nhood machinery is used (n_cls >= 2 is asserted by the caller)
        check(lib.sqb_nhood_create(ctx.handle, n_cells, 0, empty_ptr.ctypes.data, None, max(n_cls, 2), C.byref(h)))
        try:
            lab = np.ascontiguousarray(clustering, dtype=np.uint32)
            check(lib.sqb_nhood_set_base(h, lab.ctypes.data, None, 0))
            states = spawn_states(seed, int(n_perms))
            check(lib.sqb_nhood_permute_upload(h, states.ctypes.data, states.shape[0]))
            check(lib.sqb_ligrec_counts(h, data_arr.ctypes.data, n_genes, inv_counts.ctypes.data, mean_c.ctypes.data, inter.ctypes.data,
                                        inter.shape[0], cpairs.ctypes.data, cpairs.shape[0], valid_u8.ctypes.data, counts.ctypes.data))
        finally:
            lib.sqb_nhood_destroy(h)
    pvalues = counts.astype(np.float64) / n_perms
    pvalues[~valid] = np.nan
    return TempResult(means=res_means, pvalues=pvalues)


# ---------------------------------------------------------------------------------------------------------
# host glue: interactions / complexes / result frames
# ---------------------------------------------------------------------------------------------------------
def _as_interaction_frame(interactions: Any) -> pd.DataFrame:
    if interactions is None:
        raise NotImplementedError("`interactions=None` fetches the interactions from omnipath (network access); pass them explicitly.")
    if isinstance(interactions, Mapping):
        interactions = pd.DataFrame(interactions)
    if isinstance(interactions, pd.DataFrame):
        for col in (SOURCE, TARGET):
            if col not in interactions.columns:
                raise KeyError(f"Column `{col!r}` is not in `interactions`.")
        df = interactions.copy()
    elif isinstance(interactions, Iterable):
        items = tuple(interactions)
        if not len(items):
            raise ValueError("No interactions were specified.")
        if isinstance(items[0], str):
            items = list(product(items, repeat=2))
        elif len(items) == 2:
            items = tuple(zip(*items, strict=False))
        if not all(len(i) == 2 for i in items):
            raise ValueError("Not all interactions are of length `2`.")
        df = pd.DataFrame(items, columns=[SOURCE, TARGET])
    else:
        raise TypeError(f"Expected either a `pandas.DataFrame`, `dict` or `iterable`, found `{type(interactions).__name__}`")
    if df.empty:
        raise ValueError("The interactions are empty")
    return df


def _resolve_complexes(df: pd.DataFrame, data: pd.DataFrame, policy: ComplexPolicy) -> pd.DataFrame:
    """``_filter_interactions_complexes`` (``_ligrec.py:389-451``): ``'A_B_C'`` names a protein complex."""
    if policy == ComplexPolicy.MIN:
        def pick(name):
            if name is None:
                return None
            if "_" not in name:
                return name
            parts = [c for c in name.split("_") if c in data.columns]
            if not parts:
                return None
            if len(parts) == 1:
                return parts[0]
            means = data[parts].mean()
            return str(means.index[means.argmin()])  # the least expressed component, like CellPhoneDB

        df[SOURCE] = df[SOURCE].apply(pick)
        df[TARGET] = df[TARGET].apply(pick)
        return df
    if policy == ComplexPolicy.ALL:
        src = df.pop(SOURCE).apply(lambda s: str(s).split("_")).explode()
        src.name = SOURCE
        tgt = df.pop(TARGET).apply(lambda s: str(s).split("_")).explode()
        tgt.name = TARGET
        df = pd.merge(df, src, how="left", left_index=True, right_index=True)
        return pd.merge(df, tgt, how="left", left_index=True, right_index=True)
    raise NotImplementedError(f"Complex policy {policy!r} is not implemented.")


def _multipletests(p: np.ndarray, method: str, alpha: float) -> np.ndarray:
    from ._ppatterns import _multipletests as mt

    return mt(p, method)


def _fdr_correct(pvals: pd.DataFrame, corr_method: str, corr_axis: CorrAxis, alpha: float) -> pd.DataFrame:
    """``_fdr_correct`` (``_ligrec.py:57-90``): NaNs (untested pairs) count as 1 during the correction and stay NaN."""
    def fdr(col: pd.Series):
        q = _multipletests(np.nan_to_num(col.values, copy=True, nan=1.0), corr_method, alpha)
        q[np.isnan(col.values)] = np.nan
        return pd.arrays.SparseArray(q, dtype=q.dtype, fill_value=np.nan)

    if corr_axis == CorrAxis.CLUSTERS:
        pvals = pvals.apply(fdr)  # clusters are in columns
    elif corr_axis == CorrAxis.INTERACTIONS:
        pvals = pvals.T.apply(fdr).T
    else:
        raise NotImplementedError(f"FDR correction for `{corr_axis}` is not implemented.")
    return pvals


def ligrec(adata: Any, cluster_key: str, interactions: Any = None, complex_policy: str = ComplexPolicy.MIN.v, threshold: float = 0.01,
           corr_method: str | None = None, corr_axis: str = CorrAxis.CLUSTERS.v, use_raw: bool = True, copy: bool = False,
           key_added: str | None = None, gene_symbols: str | None = None, *, n_perms: int = 1000, seed: int | None = None,
           clusters: Any = None, alpha: float = 0.05, n_jobs: int | None = None, show_progress_bar: bool = True,
           table_key: str | None = None, device: int | None = None) -> Mapping[str, pd.DataFrame] | None:
    """Permutation test of receptor-ligand interactions between clusters (``_ligrec.py:543-613``); returns / writes
    ``adata.uns[f'{cluster_key}_ligrec'] = {'means', 'pvalues', 'metadata'}`` (frames indexed by (source, target), columns
    (cluster_1, cluster_2)).  ``n_jobs`` / ``show_progress_bar`` are accepted and ignored."""
    adata = extract_adata_if_sdata(adata, table_key=table_key)
    if not adata.shape[0]:
        raise ValueError("No cells are in `adata.obs_names`.")
    if not adata.shape[1]:
        raise ValueError("No genes are in `adata.var_names`.")
    src_adata = adata
    if use_raw:
        if adata.raw is None:
            raise AttributeError("No `.raw` attribute found. Try specifying `use_raw=False`.")
        src_adata = adata.raw
    var_names = src_adata.var_names if gene_symbols is None else src_adata.var[gene_symbols]
    from scipy.sparse import issparse

    X = src_adata.X
    X = X.toarray() if issparse(X) else np.asarray(X)
    data = pd.DataFrame(X, index=adata.obs_names, columns=pd.Index(var_names).astype(str)).fillna(0.0)

    complex_policy = ComplexPolicy(complex_policy)
    inter = _as_interaction_frame(interactions)
    # upper-case gene symbols, drop missing / repeated interactions and repeated genes (_ligrec.py:203-218)
    data.columns = data.columns.str.upper()
    inter[SOURCE] = inter[SOURCE].str.upper()
    inter[TARGET] = inter[TARGET].str.upper()
    inter = inter.dropna(subset=[SOURCE, TARGET], how="any").drop_duplicates(subset=[SOURCE, TARGET], keep="first")
    data = data.loc[:, ~data.columns.duplicated()]
    inter = _resolve_complexes(inter, data, complex_policy)
    inter = inter[inter[SOURCE].isin(data.columns) & inter[TARGET].isin(data.columns)]
    if inter.empty:
        raise ValueError("After filtering by genes, no interactions remain.")
    filtered = data.loc[:, list(set(inter[SOURCE]) | set(inter[TARGET]))].copy()
    inter = inter.drop_duplicates(subset=[SOURCE, TARGET], keep="first")

    # ---- test (PermutationTest.test, :230-370)
    assert_positive(n_perms, name="n_perms")
    assert_categorical_obs(adata, key=cluster_key)
    if corr_method is not None:
        corr_axis = CorrAxis(corr_axis)
    cats = adata.obs[cluster_key].cat.categories
    if len(cats) <= 1:
        raise ValueError(f"Expected at least `2` clusters, found `{len(cats)}`.")
    pairs_df = inter[[SOURCE, TARGET]]
    filtered["clusters"] = adata.obs[cluster_key].astype("string").astype("category").values
    if clusters is None:
        clusters = list(map(str, cats))
    if all(isinstance(c, str) for c in clusters):
        clusters = list(product(clusters, repeat=2))
    known = filtered["clusters"].cat.categories
    checked = []
    for needle in clusters:
        if len(needle) != 2:
            raise ValueError(f"Expected a `tuple` of length `2`, found `{len(needle)}`.")
        for c in needle:
            if c not in known:
                raise ValueError(f"Invalid cluster `{c!r}`.")
        checked.append(tuple(needle))
    clusters = sorted(checked)
    flat = list({c for cs in clusters for c in cs})
    sub = filtered.loc[np.isin(filtered["clusters"], flat), :].copy()
    sub["clusters"] = sub["clusters"].cat.remove_unused_categories()
    cat = sub["clusters"].cat
    cluster_mapper = dict(zip(cat.categories, range(len(cat.categories)), strict=False))
    gene_mapper = dict(zip(sub.columns[:-1], range(len(sub.columns) - 1), strict=False))
    sub.columns = [gene_mapper[c] if c != "clusters" else c for c in sub.columns]
    clusters_ = np.array([[cluster_mapper[a], cluster_mapper[b]] for a, b in clusters], dtype=np.uint32)
    sub["clusters"] = cat.rename_categories(cluster_mapper)
    inter_ = np.vectorize(lambda g: gene_mapper[g])(pairs_df.values)

    start = time.perf_counter()
    logg.info("Running `%d` permutations on `%d` interactions and `%d` cluster combinations on the GPU", n_perms, len(pairs_df), len(clusters))
    res_t = ligrec_analysis(sub, inter_, clusters_, threshold=threshold, n_perms=int(n_perms), seed=shared_seed(seed), ctx=default_context(device))
    index = pd.MultiIndex.from_frame(pairs_df, names=[SOURCE, TARGET])
    columns = pd.MultiIndex.from_tuples(clusters, names=["cluster_1", "cluster_2"])
    res: dict[str, pd.DataFrame] = {
        "means": pd.DataFrame({c: pd.arrays.SparseArray(res_t.means[:, i], fill_value=0) for i, c in enumerate(columns)}, index=index),
        "pvalues": pd.DataFrame({c: pd.arrays.SparseArray(res_t.pvalues[:, i], fill_value=np.nan) for i, c in enumerate(columns)}, index=index),
        "metadata": inter[inter.columns.difference([SOURCE, TARGET])],
    }
    res["metadata"].index = res["means"].index.copy()
    if corr_method is not None:
        res["pvalues"] = _fdr_correct(res["pvalues"], corr_method, corr_axis, alpha)
    if copy:
        logg.info("Finish (%.3fs)", time.perf_counter() - start)
        return res
    _save_data(adata, attr="uns", key=Key.uns.ligrec(cluster_key, key_added), data=res, time_start=start)
    return None
