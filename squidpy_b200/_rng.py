"""Per-permutation numpy generators -> flat PCG64 state table for the device.

``spawn_generators(seed, n)`` of the reference (``src/squidpy/_utils.py:240-241``) is
``[default_rng(s) for s in SeedSequence(seed).spawn(n)]``.  The device replays each generator's stream, so the
host only has to hand over the initial 128-bit state and increment of every generator.
"""

from __future__ import annotations

import numpy as np

_M64 = (1 << 64) - 1


def spawn_generators(seed: int | None, n: int) -> list[np.random.Generator]:
    return [np.random.default_rng(s) for s in np.random.SeedSequence(seed).spawn(n)]


def generator_state(gen: np.random.Generator) -> tuple[int, int, int, int, int, int]:
    st = gen.bit_generator.state
    if st["bit_generator"] != "PCG64":
        raise TypeError(f"Expected a PCG64 generator, found `{st['bit_generator']}`.")
    s, inc = st["state"]["state"], st["state"]["inc"]
    return (s >> 64, s & _M64, inc >> 64, inc & _M64, int(st["has_uint32"]), int(st["uinteger"]))


_INIT_A, _MULT_A = 0x43B0D7E5, 0x931E8875
_INIT_B, _MULT_B = 0x8B51F9DD, 0x58F38DED
_MIX_L, _MIX_R = 0xCA01F9DD, 0x4973F715
_PCG_MULT = 0x2360ED051FC65DA44385DF649FCCF645
_M128 = (1 << 128) - 1


def _uint32_words(x) -> list[int]:
    """numpy's ``_coerce_to_uint32_array`` for the cases SeedSequence entropy takes here: non-negative ints (split into
    little-endian 32-bit words) and (nested) sequences of them."""
    if isinstance(x, (int, np.integer)):
        x = int(x)
        if x < 0:
            raise ValueError("expected non-negative integer")
        words = [x & 0xFFFFFFFF]
        x >>= 32
        while x:
            words.append(x & 0xFFFFFFFF)
            x >>= 32
        return words
    out: list[int] = []
    for v in x:
        out.extend(_uint32_words(v))
    return out


def _spawn_states_vectorised(root: np.random.SeedSequence, start: int, stop: int) -> np.ndarray:
    """SeedSequence(entropy, spawn_key=(i,)) -> PCG64 state for i in [start, stop), all children at once with uint32
    array arithmetic (numpy/random/bit_generator.pyx: mix_entropy / generate_state; _pcg64.pyx: pcg64_srandom_r).
    A Python loop over 1000 SeedSequence + default_rng objects costs ~11 ms, this ~1 ms."""
    m = stop - start
    pool_size = root.pool_size
    run = _uint32_words(root.entropy)
    key = [w for k in root.spawn_key for w in _uint32_words(k)]
    if len(run) < pool_size:  # a non-empty spawn key follows: pad the run entropy to the pool size
        run = run + [0] * (pool_size - len(run))
    idx = np.arange(start, stop, dtype=np.uint64)
    # child index words (little endian 32-bit): one word below 2^32, two above
    child_words = [(idx & np.uint64(0xFFFFFFFF)).astype(np.uint32)]
    if stop > (1 << 32):
        raise ValueError("more than 2^32 children are not supported")
    ent = [np.full(m, w, dtype=np.uint32) for w in run + key] + child_words
    u32 = np.uint32
    hc = _INIT_A  # the hash constant evolves identically for every child: keep it a Python int

    def hashmix(v):
        nonlocal hc
        v = v ^ u32(hc)
        hc = (hc * _MULT_A) & 0xFFFFFFFF
        v = v * u32(hc)
        return v ^ (v >> u32(16))

    def mix(x, y):
        r = u32(_MIX_L) * x - u32(_MIX_R) * y
        return r ^ (r >> u32(16))

    with np.errstate(over="ignore"):
        pool = [hashmix(ent[i] if i < len(ent) else np.zeros(m, u32)) for i in range(pool_size)]
        for i_src in range(pool_size):
            for i_dst in range(pool_size):
                if i_src != i_dst:
                    pool[i_dst] = mix(pool[i_dst], hashmix(pool[i_src]))
        for i_src in range(pool_size, len(ent)):
            for i_dst in range(pool_size):
                pool[i_dst] = mix(pool[i_dst], hashmix(ent[i_src]))
        # generate_state(4, uint64) = 8 uint32 words
        hb = _INIT_B
        words = []
        for i in range(8):
            v = pool[i % pool_size] ^ u32(hb)
            hb = (hb * _MULT_B) & 0xFFFFFFFF
            v = v * u32(hb)
            words.append(v ^ (v >> u32(16)))
    val = [(words[2 * k].astype(np.uint64) | (words[2 * k + 1].astype(np.uint64) << np.uint64(32))) for k in range(4)]
    out = np.zeros((m, 6), dtype=np.uint64)
    # pcg64_srandom_r: inc = (initseq << 1) | 1; state = 0; step; state += initstate; step   (128-bit, per child)
    v0, v1, v2, v3 = (v.tolist() for v in val)
    for k in range(m):
        inc = ((((v2[k] << 64) | v3[k]) << 1) | 1) & _M128
        st = inc  # 0 * MULT + inc
        st = (st + ((v0[k] << 64) | v1[k])) & _M128
        st = (st * _PCG_MULT + inc) & _M128
        out[k, 0] = st >> 64
        out[k, 1] = st & _M64
        out[k, 2] = inc >> 64
        out[k, 3] = inc & _M64
    return out


def spawn_states(seed: int | None, n: int, start: int = 0, stop: int | None = None) -> np.ndarray:
    """(stop-start, 6) uint64 table {state_hi, state_lo, inc_hi, inc_lo, has_uint32, uinteger} of generators
    ``start..stop`` out of the ``n`` children of ``SeedSequence(seed)`` (child ``i`` has ``spawn_key=(i,)``, so a
    shard can build only its own generators)."""
    stop = n if stop is None else stop
    root = np.random.SeedSequence(seed)
    try:
        return _spawn_states_vectorised(root, start, stop)
    except (TypeError, ValueError):  # exotic entropy types: let numpy build every generator
        pass
    out = np.empty((stop - start, 6), dtype=np.uint64)
    for k, i in enumerate(range(start, stop)):
        child = np.random.SeedSequence(root.entropy, spawn_key=root.spawn_key + (i,), pool_size=root.pool_size)
        out[k] = generator_state(np.random.default_rng(child))
    return out
