"""Stub-import loader for the UNMODIFIED reference hot-path modules (test infrastructure only).

This file is part of ``oracle/`` — test infrastructure, never imported by the product package.
It executes the reference's own ``gr/_nhood.py``, ``gr/_ppatterns.py``, ``gr/_ripley.py`` and
``gr/neighbors.py`` straight from ``/root/reference/src`` (read-only, not copied) by pre-registering
light stub modules for the heavyweight dependencies that are absent from this image (anndata, scanpy,
spatialdata, rustworkx, docrep, xarray, statsmodels ...), so that none of squidpy's package
``__init__`` files run.  It exists so that (a) the CPU restatements in ``oracle/`` can be pinned
against the real reference code in the build container and (b) ``tests/golden/make_golden.py`` can
generate the committed golden vectors.  ``/root/reference`` does not exist on the GPU box: everything
that calls :func:`load` must skip cleanly when :func:`available` is False.
"""

from __future__ import annotations

import importlib
import importlib.metadata as _md
import os
import sys
import types

REF_SRC = os.environ.get("SQUIDPY_REF", "/root/reference/src")
_loaded: dict[str, types.ModuleType] | None = None


def available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "squidpy", "gr"))


def _stub(name: str, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # type: ignore[attr-defined]
    sys.modules[name] = m
    return m


class _Any:
    def __init__(self, *a, **k):
        pass


class _Logg:
    @staticmethod
    def info(*a, **k):
        return None

    warning = hint = debug = error = info


class _DocstringProcessor:
    def __init__(self, **k):
        pass

    def dedent(self, f):
        return f

    def get_sections(self, **k):
        return lambda f: f

    get_full_description = get_sections


def load(morans_i=None, gearys_c=None, multipletests=None) -> dict[str, types.ModuleType]:
    """Import the reference hot-path modules.  ``morans_i``/``gearys_c`` (scanpy is not installed and its
    source is not under /root/reference) and ``multipletests`` (statsmodels) are bound to the callables given."""
    global _loaded
    if _loaded is not None:
        if morans_i is not None:
            _loaded["pp"].morans_i = morans_i
        if gearys_c is not None:
            _loaded["pp"].gearys_c = gearys_c
        if multipletests is not None:
            _loaded["pp"].multipletests = multipletests
        return _loaded
    if not available():
        raise RuntimeError(f"reference sources not found under {REF_SRC}")
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_sqb_oracle")

    _stub("anndata", AnnData=type("AnnData", (_Any,), {}), OldFormatWarning=Warning)
    _stub("anndata.utils", make_index_unique=lambda x: x)
    _stub("anndata._core")
    _stub(
        "anndata._core.views",
        ArrayView=type("ArrayView", (), {}),
        SparseCSRMatrixView=type("SparseCSRMatrixView", (), {}),
        SparseCSCMatrixView=type("SparseCSCMatrixView", (), {}),
    )
    _stub("scanpy", logging=_Logg)
    _stub("scanpy.logging", info=_Logg.info, warning=_Logg.info)
    _stub("scanpy.get", obs_df=None)
    _stub("scanpy.metrics", morans_i=morans_i, gearys_c=gearys_c)
    for n in (
        "scanpy.plotting",
        "scanpy.plotting.legacy",
        "scanpy.plotting.legacy._tools",
        "scanpy.plotting.legacy._tools.scatterplots",
        "scanpy.plotting.legacy._utils",
        "scanpy.plotting.legacy.palettes",
        "scanpy.plotting.legacy.mpl_settings",
    ):
        _stub(
            n,
            _add_categorical_legend=None,
            _panel_grid=None,
            add_colors_for_categorical_sample_annotation=None,
            set_default_colors_for_categorical_obs=None,
            default_102=None,
            FRAMEON=True,
            VECTOR_FRIENDLY=True,
        )
    sys.modules["scanpy.plotting.legacy"].mpl_settings = sys.modules["scanpy.plotting.legacy.mpl_settings"]
    _stub("spatialdata", SpatialData=type("SpatialData", (_Any,), {}))
    _stub("spatialdata.models", Image2DModel=_Any, Labels2DModel=_Any)
    _stub("rustworkx")
    _stub("xarray", DataArray=_Any)
    _stub("docrep", DocstringProcessor=_DocstringProcessor)
    _stub("statsmodels")
    _stub("statsmodels.stats")
    _stub("statsmodels.stats.multitest", multipletests=multipletests)
    _stub("fast_array_utils", stats=types.SimpleNamespace(sum=lambda a, axis=None: a.sum(axis=axis)))
    _orig = _md.version
    _md.version = lambda n: "0.12.0" if n == "anndata" else _orig(n)
    for pkg in ("squidpy", "squidpy.gr", "squidpy._constants"):
        m = types.ModuleType(pkg)
        m.__path__ = [f"{REF_SRC}/{pkg.replace('.', '/')}"]  # type: ignore[attr-defined]
        sys.modules[pkg] = m
    _loaded = {
        "nh": importlib.import_module("squidpy.gr._nhood"),
        "pp": importlib.import_module("squidpy.gr._ppatterns"),
        "rp": importlib.import_module("squidpy.gr._ripley"),
        "nb": importlib.import_module("squidpy.gr.neighbors"),
        "utils": importlib.import_module("squidpy._utils"),
        "grutils": importlib.import_module("squidpy.gr._utils"),
    }
    return _loaded
