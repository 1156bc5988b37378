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


def spawn_states(seed: int | None, n: int, start: int = 0, stop: int | None = None) -> np.ndarray:
    """(stop-start, 6) uint64 table {state_hi, state_lo, inc_hi, inc_lo, has_uint32, uinteger} of generators
    ``start..stop`` out of the ``n`` children of ``SeedSequence(seed)`` (child ``i`` has ``spawn_key=(i,)``, so a
    shard can build only its own generators)."""
    stop = n if stop is None else stop
    root = np.random.SeedSequence(seed)
    out = np.empty((stop - start, 6), dtype=np.uint64)
    for k, i in enumerate(range(start, stop)):
        child = np.random.SeedSequence(root.entropy, spawn_key=root.spawn_key + (i,), pool_size=root.pool_size)
        out[k] = generator_state(np.random.default_rng(child))
    return out
