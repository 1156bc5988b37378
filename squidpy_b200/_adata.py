"""Minimal duck-typed stand-in for :class:`anndata.AnnData` (``anndata`` is not installed in this image).

The ``gr`` functions only touch ``.obs[key]`` (pandas categorical), ``.obsp[key]`` (scipy CSR), ``.obsm[key]``,
``.uns``, ``.X``/``.layers``/``.raw``, ``.var``, ``.var_names``, ``.shape`` and ``adata[:, genes]`` — exactly the
AnnData surface the reference uses on this path (``src/squidpy/gr/_utils.py:25-86``,
``src/squidpy/gr/_ppatterns.py:154-194``) — so a real AnnData works unchanged; this class exists for the tests,
the benchmark and users without anndata.
"""

from __future__ import annotations

from typing import Any

import numpy as np
import pandas as pd


class AnnDataLite:
    def __init__(self, X=None, obs: pd.DataFrame | None = None, var: pd.DataFrame | None = None, obsm=None, obsp=None,
                 uns=None, layers=None, raw=None, shape: tuple[int, int] | None = None):
        if X is not None:
            shape = X.shape
        elif shape is None:
            n_obs = len(obs) if obs is not None else 0
            n_var = len(var) if var is not None else 0
            shape = (n_obs, n_var)
        self.X = X
        self._shape = tuple(shape)
        self.obs = obs if obs is not None else pd.DataFrame(index=pd.RangeIndex(self._shape[0]).astype(str))
        self.var = var if var is not None else pd.DataFrame(index=pd.RangeIndex(self._shape[1]).astype(str))
        self.obsm = dict(obsm or {})
        self.obsp = dict(obsp or {})
        self.uns = dict(uns or {})
        self.layers = dict(layers or {})
        self.raw = raw

    @property
    def shape(self) -> tuple[int, int]:
        return self._shape

    @property
    def n_obs(self) -> int:
        return self._shape[0]

    @property
    def n_vars(self) -> int:
        return self._shape[1]

    @property
    def var_names(self) -> pd.Index:
        return self.var.index

    @property
    def obs_names(self) -> pd.Index:
        return self.obs.index

    def copy(self) -> "AnnDataLite":
        import copy as _copy

        return AnnDataLite(
            X=None if self.X is None else self.X.copy(), obs=self.obs.copy(), var=self.var.copy(),
            obsm={k: v.copy() for k, v in self.obsm.items()}, obsp={k: v.copy() for k, v in self.obsp.items()},
            uns=_copy.deepcopy(self.uns), layers={k: v.copy() for k, v in self.layers.items()}, raw=self.raw,
            shape=self._shape,
        )

    def _var_indexer(self, sel: Any) -> np.ndarray:
        if isinstance(sel, slice):
            return np.arange(self.n_vars)[sel]
        if isinstance(sel, (pd.Series, pd.Index)):
            sel = sel.to_numpy()
        sel = np.asarray(sel)
        if sel.dtype == bool:
            return np.flatnonzero(sel)
        if sel.dtype.kind in "iu":
            return sel.astype(np.int64)
        ix = self.var.index.get_indexer(np.atleast_1d(sel))
        if (ix < 0).any():
            missing = [str(s) for s, i in zip(np.atleast_1d(sel), ix) if i < 0]
            raise KeyError(f"Values {missing} are not valid var names or indices.")
        return ix

    def __getitem__(self, index):
        if not isinstance(index, tuple) or len(index) != 2:
            raise NotImplementedError("AnnDataLite only supports `adata[:, vars]` / `adata[obs, vars]` indexing")
        rows, cols = index
        ci = self._var_indexer(cols)
        ri = slice(None) if (isinstance(rows, slice) and rows == slice(None)) else rows
        X = None if self.X is None else self.X[ri][:, ci]
        layers = {k: v[ri][:, ci] for k, v in self.layers.items()}
        obs = self.obs if isinstance(ri, slice) else self.obs.iloc[ri]
        return AnnDataLite(X=X, obs=obs, var=self.var.iloc[ci], obsm=self.obsm if isinstance(ri, slice) else {},
                           obsp=self.obsp if isinstance(ri, slice) else {}, uns=self.uns, layers=layers,
                           shape=(len(obs), len(ci)))
