"""The CTA-parallel shuffle ALGORITHM (tests/emu_shuffle.py mirrors nhood_shuffle_cta_kernel step by step) against
numpy's Generator.shuffle, over sizes around powers of two, several thread counts and library segments."""

from __future__ import annotations

import random

import numpy as np
import pytest

from oracle import ref
from tests.emu_shuffle import emu_shuffle


def _state(seed):
    st = ref.spawn_states(seed, 1)[0]
    return st, (int(st[0]) << 64) | int(st[1]), (int(st[2]) << 64) | int(st[3])


@pytest.mark.parametrize("nt", [4, 32, 128])
def test_emulated_cta_shuffle_matches_numpy(nt):
    rnd = random.Random(nt)
    for n in [1, 2, 3, 4, 5, 8, 9, 16, 17, 33, 64, 65, 127, 128, 129, 255, 257, 1000, 1024, 1025, 4099]:
        seed = rnd.randrange(10**6)
        st, state, inc = _state(seed)
        gen = np.random.default_rng(np.random.SeedSequence(seed).spawn(1)[0])
        exp = np.arange(n, dtype=np.uint32)
        gen.shuffle(exp)
        got = emu_shuffle(list(range(n)), state, inc, [(0, n)], NT=nt, rng=rnd)
        np.testing.assert_array_equal(np.array(got, dtype=np.uint32), exp)


def test_emulated_cta_shuffle_segments():
    rnd = random.Random(7)
    for n in [10, 100, 513, 3000]:
        seed = rnd.randrange(10**6)
        st, state, inc = _state(seed)
        cuts = sorted(rnd.sample(range(1, n), 3))
        bounds = [0] + cuts + [n]
        segs = [(bounds[k], bounds[k + 1] - bounds[k]) for k in range(4)]
        exp = np.arange(n, dtype=np.uint32)
        s6 = st.copy()
        for b, m in segs:
            exp[b : b + m] = ref.shuffle_u32(s6, exp[b : b + m])
        got = emu_shuffle(list(range(n)), state, inc, segs, NT=16, rng=rnd)
        np.testing.assert_array_equal(np.array(got, dtype=np.uint32), exp)
