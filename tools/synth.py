"""Synthetic inputs for tests and benchmarks.  Pure numpy/scipy; no reference imports."""

from __future__ import annotations

import numpy as np
import pandas as pd
import scipy.sparse as sp


def hex_coords(rows: int, cols: int, scale: float = 100.0) -> np.ndarray:
    """Visium-like hexagonal lattice, row-major: x = col + 0.5*(row%2), y = row*sqrt(3)/2."""
    r, c = np.divmod(np.arange(rows * cols), cols)
    return np.stack([(c + 0.5 * (r % 2)) * scale, r * (np.sqrt(3.0) / 2.0) * scale], axis=1)


def hex_graph(rows: int, cols: int) -> sp.csr_matrix:
    """6-neighbour adjacency of the lattice above (what the reference's GridBuilder(n_neighs=6) yields for it —
    checked in tests/test_oracle_vs_reference.py): CSR float32 ones, int32 sorted indices, both directions stored."""
    n = rows * cols
    r, c = np.divmod(np.arange(n, dtype=np.int64), cols)
    shift = np.where(r % 2 == 1, 1, -1)
    cand = [(r, c - 1), (r, c + 1), (r - 1, c), (r + 1, c), (r - 1, c + shift), (r + 1, c + shift)]
    src, dst = [], []
    for rr, cc in cand:
        ok = (rr >= 0) & (rr < rows) & (cc >= 0) & (cc < cols)
        src.append(np.flatnonzero(ok))
        dst.append((rr[ok] * cols + cc[ok]))
    src, dst = np.concatenate(src), np.concatenate(dst)
    a = sp.csr_matrix((np.ones(src.size, np.float32), (src, dst)), shape=(n, n))
    a.sort_indices()
    a.indices = a.indices.astype(np.int32)
    a.indptr = a.indptr.astype(np.int32)
    return a


def knn_graph(coords: np.ndarray, k: int = 6) -> sp.csr_matrix:
    """Directed exact kNN graph (reference KNNBuilder contract: float32 ones, int32 indices, not symmetric)."""
    from sklearn.neighbors import NearestNeighbors

    n = coords.shape[0]
    _, col = NearestNeighbors(n_neighbors=k, metric="euclidean").fit(coords).kneighbors()
    row = np.repeat(np.arange(n), k)
    a = sp.csr_matrix((np.ones(n * k, np.float32), (row, col.reshape(-1))), shape=(n, n))
    a.sort_indices()
    a.indices = a.indices.astype(np.int32)
    a.indptr = a.indptr.astype(np.int32)
    return a


def categorical_labels(n: int, n_cls: int, seed: int = 0, prefix: str = "c") -> pd.Series:
    codes = np.random.default_rng(seed).integers(0, n_cls, n)
    cats = [f"{prefix}{i:03d}" for i in range(n_cls)]
    return pd.Series(pd.Categorical.from_codes(codes, categories=cats))


def make_adata(coords: np.ndarray, graph: sp.csr_matrix | None, labels: pd.Series | None, cluster_key: str = "cluster", X=None, var_names=None):
    from squidpy_b200._adata import AnnDataLite

    n = coords.shape[0]
    obs = pd.DataFrame(index=pd.RangeIndex(n).astype(str))
    if labels is not None:
        obs[cluster_key] = labels.values
    var = None
    if X is not None:
        var = pd.DataFrame(index=pd.Index(var_names if var_names is not None else [f"g{i}" for i in range(X.shape[1])]))
    ad = AnnDataLite(X=X, obs=obs, var=var, obsm={"spatial": coords}, shape=(n, 0 if X is None else X.shape[1]))
    if graph is not None:
        ad.obsp["spatial_connectivities"] = graph
    return ad


def expression_csr(n_obs: int, n_genes: int, density: float = 0.1, smooth_frac: float = 0.01, coords: np.ndarray | None = None,
                   seed: int = 0, dtype=np.float32) -> sp.csr_matrix:
    """cells x genes CSR float32 with exactly round(n_genes*density) sorted, distinct non-zeros per cell (log1p of
    Poisson counts).  Gene columns are split into strides; a cell expresses one gene per stride.  In a fraction
    `smooth_frac` of the strides the expressed gene is chosen by a smooth spatial field, so those genes live in
    contiguous spatial bands and their Moran's I spans ~0.1..0.9; everywhere else it is iid (I ~ E[I] ~ 0)."""
    rng = np.random.default_rng(seed)
    nnz_per_row = max(1, int(round(n_genes * density)))
    indptr = np.arange(0, (n_obs + 1) * nnz_per_row, nnz_per_row, dtype=np.int64)
    base = (np.arange(nnz_per_row, dtype=np.int64) * n_genes) // nnz_per_row
    width = max(1, n_genes // nnz_per_row)
    off = rng.integers(0, width, size=(n_obs, nnz_per_row), dtype=np.int64)
    if coords is not None and smooth_frac > 0 and width > 1:
        n_s = max(1, int(round(nnz_per_row * smooth_frac)))
        strides = rng.choice(nnz_per_row, n_s, replace=False)
        xy = (coords - coords.min(0)) / (np.ptp(coords, axis=0) + 1e-9)
        for q, k in enumerate(strides):
            fx, fy, ph = 1 + q % 3, 1 + (q // 3) % 3, rng.random() * 6.28
            field = 0.5 + 0.5 * np.sin(2 * np.pi * (fx * xy[:, 0] + 0.5 * fy * xy[:, 1]) + ph) * np.cos(np.pi * fy * xy[:, 1])
            off[:, k] = np.minimum((field * width).astype(np.int64), width - 1)
    indices = np.minimum(base[None, :] + off, n_genes - 1)
    vals = np.log1p(rng.poisson(1.5, size=(n_obs, nnz_per_row)) + 1.0).astype(dtype)
    x = sp.csr_matrix((vals.reshape(-1), indices.reshape(-1).astype(np.int32), indptr), shape=(n_obs, n_genes))
    x.has_sorted_indices = True
    return x


def thomas_points(n: int, n_parents: int = 200, sigma: float = 150.0, extent: float = 1.0e4, seed: int = 0) -> np.ndarray:
    """Mildly clustered planar point pattern (Thomas process) in [0, extent]^2, float64."""
    rng = np.random.default_rng(seed)
    parents = rng.random((n_parents, 2)) * extent
    which = rng.integers(0, n_parents, n)
    pts = parents[which] + rng.normal(0.0, sigma, (n, 2))
    return np.clip(pts, 0.0, extent)


def dirichlet_labels(n: int, n_cls: int, seed: int = 0, alpha: float = 1.0) -> pd.Series:
    rng = np.random.default_rng(seed)
    p = rng.dirichlet(np.full(n_cls, alpha))
    codes = rng.choice(n_cls, size=n, p=p)
    codes[:n_cls] = np.arange(n_cls)  # every category present
    return pd.Series(pd.Categorical.from_codes(codes, categories=[f"k{i:02d}" for i in range(n_cls)]))
