"""Executable specification (numpy) of the fast RNG mode of nhood_enrichment (``rng="philox"``,
``sqb_nhood_permute_upload_philox``; device code: ``nhood_philox_labels_kernel`` in ``squidpy_b200/csrc/nhood.cu``).

Permutation ``p`` (global index) of library segment ``s`` (``m`` observations, in ``np.where(libraries == c)[0]`` order
like ``gr/_utils.py:208-209``) assigns to the observation at position ``r`` of the segment the label
``sorted_labels_of_segment[pi(r)]`` where ``pi`` is a 4-round Feistel network on ``2k`` bits (``k = ceil(bits(m-1)/2)``,
at least 1) with cycle walking, round keys ``philox_key(seed, p, s, round)``.
TEST INFRASTRUCTURE: the CUDA kernel is checked bit-for-bit against this file.
"""

from __future__ import annotations

import numpy as np

M64 = (1 << 64) - 1


def philox_key(seed: int, perm: int, seg: int, rnd: int) -> int:
    z = (seed + 0x9E3779B97F4A7C15 * (perm + 1)) & M64
    z ^= ((seg * 4 + rnd + 1) * 0xD1B54A32D192ED03) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    z ^= z >> 31
    return z >> 32


def feistel_perm(m: int, keys: list[int]) -> np.ndarray:
    """pi(r) for r in [0, m) (vectorised cycle walking)."""
    if m <= 1:
        return np.zeros(m, dtype=np.int64)
    bits = int(m - 1).bit_length()
    k = max(1, (bits + 1) // 2)
    mask = np.uint32((1 << k) - 1)
    x = np.arange(m, dtype=np.uint32)
    todo = np.ones(m, dtype=bool)
    with np.errstate(over="ignore"):
        while todo.any():
            v = x[todo]
            left, right = v >> np.uint32(k), v & mask
            for key in keys:
                t = (right ^ np.uint32(key)) * np.uint32(0x85EBCA6B)
                t ^= t >> np.uint32(13)
                t *= np.uint32(0xC2B2AE35)
                t ^= t >> np.uint32(16)
                left, right = right, left ^ (t & mask)
            v = (left << np.uint32(k)) | right
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
            pi = feistel_perm(idx.size, [philox_key(seed, int(p), s, r) for r in range(4)])
            out[row, idx] = srt[pi]
    return out
