"""Executable specification (numpy) of the fast RNG mode of nhood_enrichment (``rng="philox"``,
``sqb_nhood_permute_upload_philox``; device code: ``nhood_philox_labels_kernel`` in ``squidpy_b200/csrc/nhood.cu``).

Permutation ``p`` (global index) of library segment ``s`` (``m`` observations, in ``np.where(libraries == c)[0]`` order
like ``gr/_utils.py:208-209``) assigns to the observation at position ``r`` of the segment the label
``sorted_labels_of_segment[pi(r)]`` where ``pi`` is a 6-round generalised Feistel network (FE2) on
``[0, a) x [0, b)``, ``a = ceil(sqrt(m))``, ``b = ceil(m / a)``, with cycle walking; round keys ``k0 + j*k1`` with
``(k0, k1 | 1) = philox_key(seed, p, s, 0 / 1)``.
TEST INFRASTRUCTURE: the CUDA kernel is checked bit-for-bit against this file.
"""

from __future__ import annotations

import math

import numpy as np

M64 = (1 << 64) - 1


def philox_key(seed: int, perm: int, seg: int, which: int) -> int:
    z = (seed + 0x9E3779B97F4A7C15 * (perm + 1)) & M64
    z ^= ((seg * 2 + which + 1) * 0xD1B54A32D192ED03) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    z ^= z >> 31
    return z >> 32


def _hash(v: np.ndarray) -> np.ndarray:
    v = v * np.uint32(0x9E3779B1)
    v ^= v >> np.uint32(15)
    v *= np.uint32(0x85EBCA77)
    return v


def radices(m: int) -> tuple[int, int]:
    a = math.isqrt(m)
    if a * a < m:
        a += 1
    a = max(a, 1)
    return a, max(1, -(-m // a))


def feistel_perm(m: int, k0: int, k1: int) -> np.ndarray:
    """pi(r) for r in [0, m) (vectorised cycle walking)."""
    if m <= 1:
        return np.zeros(m, dtype=np.int64)
    a, b = radices(m)
    x = np.arange(m, dtype=np.uint32)
    todo = np.ones(m, dtype=bool)
    k1 |= 1
    with np.errstate(over="ignore"):
        while todo.any():
            v = x[todo]
            left, right = v // np.uint32(b), v % np.uint32(b)
            for j in range(6):
                kj = np.uint32((k0 + j * k1) & 0xFFFFFFFF)
                if j % 2 == 0:
                    add = ((_hash(right ^ kj).astype(np.uint64) * np.uint64(a)) >> np.uint64(32)).astype(np.uint32)
                    left = (left + add) % np.uint32(a)
                else:
                    add = ((_hash(left ^ kj).astype(np.uint64) * np.uint64(b)) >> np.uint64(32)).astype(np.uint32)
                    right = (right + add) % np.uint32(b)
            v = left * np.uint32(b) + right
            x[todo] = v
            todo[todo] = v >= m
    return x.astype(np.int64)


def philox_labels(base: np.ndarray, seed: int, perms, lib_codes=None, n_libs: int = 0) -> np.ndarray:
    """uint32 (len(perms), n) label vectors in the original observation order."""
    base = np.asarray(base)
    n = base.size
    if lib_codes is None:
        groups = [np.arange(n)]
    else:
        groups = [np.where(np.asarray(lib_codes) == c)[0] for c in range(n_libs)]
    out = np.empty((len(perms), n), dtype=np.uint32)
    for row, p in enumerate(perms):
        for s, idx in enumerate(groups):
            if idx.size == 0:
                continue
            srt = np.sort(base[idx])
            pi = feistel_perm(idx.size, philox_key(seed, int(p), s, 0), philox_key(seed, int(p), s, 1))
            out[row, idx] = srt[pi]
    return out
