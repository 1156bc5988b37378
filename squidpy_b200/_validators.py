"""Argument checks with the reference's exception types and messages
(``src/squidpy/_validators.py:68-70,96-112``, ``src/squidpy/gr/_utils.py:25-74``)."""

from __future__ import annotations

from typing import Any

from pandas import CategoricalDtype
from pandas.api.types import infer_dtype


def assert_positive(value: float, *, name: str) -> None:
    if value <= 0:
        raise ValueError(f"Expected `{name}` to be positive, found `{value}`.")


def assert_key_in_adata(adata: Any, key: str, *, attr: str, extra_msg: str = "") -> None:
    container = getattr(adata, attr)
    if key not in container:
        available = list(container.keys()) if hasattr(container, "keys") else list(container)
        msg = f"Key `{key!r}` not found in `adata.{attr}`. Available keys: {available}."
        if extra_msg:
            msg = f"{msg} {extra_msg}"
        raise KeyError(msg)


def extract_adata_if_sdata(adata: Any, *, table_key: str | None = None) -> Any:
    """``SpatialData`` -> its table; anything else (AnnData or a duck-typed equivalent) is returned as is."""
    tables = getattr(adata, "tables", None)
    if tables is not None and not hasattr(adata, "obsp"):
        if table_key is None:
            raise TypeError("missing required keyword-only argument: 'table_key'")
        if table_key not in tables:
            raise ValueError(f"Table {table_key!r} not found in SpatialData. Available tables: {list(tables.keys())}")
        return tables[table_key]
    return adata


def assert_categorical_obs(adata: Any, key: str) -> None:
    if key not in adata.obs:
        raise KeyError(f"Cluster key `{key}` not found in `adata.obs`.")
    if not isinstance(adata.obs[key].dtype, CategoricalDtype):
        raise TypeError(f"Expected `adata.obs[{key!r}]` to be `categorical`, found `{infer_dtype(adata.obs[key])}`.")


def assert_connectivity_key(adata: Any, key: str) -> None:
    if key not in adata.obsp:
        key_added = key.replace("_connectivities", "")
        raise KeyError(
            f"Spatial connectivity key `{key}` not found in `adata.obsp`. "
            f"Please run `squidpy.gr.spatial_neighbors(..., key_added={key_added!r})` first."
        )


def assert_spatial_basis(adata: Any, key: str) -> None:
    if key not in adata.obsm:
        raise KeyError(f"Spatial basis `{key}` not found in `adata.obsm`.")
