from __future__ import annotations

import os
import sys

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    """GPU tests fail loudly (not skip) on a box that has a GPU; without one they are deselected by `-m "not gpu"`;
    if someone runs them without a GPU they are skipped with a clear reason."""
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dummy():
    return dict(np.load(os.path.join(GOLDEN, "dummy_adata.npz"), allow_pickle=False))


@pytest.fixture(scope="session")
def golden_cfg1():
    return dict(np.load(os.path.join(GOLDEN, "cfg1_visium5k.npz"), allow_pickle=False))


@pytest.fixture(scope="session")
def golden_pairs():
    return dict(np.load(os.path.join(GOLDEN, "pairs_jitter.npz"), allow_pickle=False))


def make_dummy_adata(g):
    """The reference's ``dummy_adata`` recipe (tests/conftest.py:109-117 there) rebuilt from the golden file."""
    import scipy.sparse as sp

    from squidpy_b200 import AnnDataLite

    adj = sp.csr_matrix((g["adj_data"], g["adj_indices"], g["adj_indptr"]), shape=(200, 200))
    obs = pd.DataFrame(
        {
            "cluster": pd.Categorical.from_codes(g["cl"], categories=["0", "1", "2"]),
            "library": pd.Categorical(["A"] * 100 + ["B"] * 100),
            "cont": g["X"][:, 0],
        },
        index=[str(i) for i in range(200)],
    )
    var = pd.DataFrame(index=[f"g{i}" for i in range(100)])
    return AnnDataLite(X=g["X"].copy(), obs=obs, var=var, obsm={"spatial": g["xy"].copy()}, obsp={"spatial_connectivities": adj})


@pytest.fixture()
def dummy_adata(golden_dummy):
    return make_dummy_adata(golden_dummy)
