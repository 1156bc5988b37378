"""CPU checks of the fast-RNG specification (tests/philox_ref.py): bijectivity, label multiset preservation, independence
of the sharding, and first-order uniformity."""

from __future__ import annotations

import numpy as np
import pytest

from tests import philox_ref as pr


@pytest.mark.parametrize("m", [1, 2, 3, 4, 5, 17, 64, 65, 1000, 4097, 70001])
def test_feistel_is_a_bijection(m):
    pi = pr.feistel_perm(m, pr.philox_key(12345, 7, 0, 0), pr.philox_key(12345, 7, 0, 1))
    a, b = pr.radices(m)
    assert a * b >= m and (a - 1) * (a - 1) < m and a * b - m < a
    assert pi.shape == (m,) and np.array_equal(np.sort(pi), np.arange(m))


def test_labels_preserve_multiset_per_library_and_differ_between_permutations():
    rng = np.random.default_rng(0)
    base = rng.integers(0, 7, 5000).astype(np.uint32)
    libs = rng.integers(0, 3, 5000)
    lab = pr.philox_labels(base, 99, range(6), libs, 3)
    for p in range(6):
        for c in range(3):
            assert np.array_equal(np.sort(lab[p, libs == c]), np.sort(base[libs == c]))
    assert len({lab[p].tobytes() for p in range(6)}) == 6
    assert np.array_equal(pr.philox_labels(base, 99, [4], libs, 3)[0], lab[4])  # permutation p does not depend on the batch


def test_first_order_uniformity():
    """Every position receives every class with the class frequency (chi-square over 2000 permutations, 40 positions)."""
    base = np.repeat(np.arange(4, dtype=np.uint32), [50, 100, 150, 200])
    lab = pr.philox_labels(base, 2024, range(2000))[:, ::12][:, :40]
    exp = np.array([50, 100, 150, 200]) / 500 * 2000
    chi = np.array([((np.bincount(lab[:, k], minlength=4) - exp) ** 2 / exp).sum() for k in range(40)])
    assert chi.mean() < 4.5 and chi.max() < 25  # 3 degrees of freedom: mean 3, P(chi2 > 25) ~ 1.5e-5


def test_second_order_statistics_match_numpy_shuffle():
    """Same-class neighbour pairs of a lattice under the bijection vs under numpy's shuffle (what nhood_enrichment
    measures).  A 4-round network fails this (mean +3.5 %, std x2); the 6-round one must match."""
    from tools import synth

    g = synth.hex_graph(71, 71)
    n = g.shape[0]
    base = np.random.default_rng(0).integers(0, 10, n).astype(np.uint32)
    rows, cols = np.repeat(np.arange(n), np.diff(g.indptr)), g.indices
    P = 300
    fast = pr.philox_labels(base, 42, range(P))
    rng = np.random.default_rng(1)
    d_fast = np.array([(lab[rows] == lab[cols]).sum() for lab in fast], dtype=np.float64)
    d_np = np.array([(lab[rows] == lab[cols]).sum() for lab in (rng.permutation(base) for _ in range(P))], dtype=np.float64)
    sd = d_np.std()
    assert abs(d_fast.mean() - d_np.mean()) < 4 * sd * np.sqrt(2 / P)
    assert abs(d_fast.std() - sd) / sd < 4 / np.sqrt(P)
