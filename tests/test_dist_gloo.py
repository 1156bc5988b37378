"""world_size-2 gloo tests of the multi-GPU host logic (sharding + the single collective per call)."""

from __future__ import annotations

import os
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent(
    """
    import os, sys, numpy as np
    sys.path.insert(0, os.environ["SQB_ROOT"])
    import torch.distributed as dist
    dist.init_process_group("gloo")
    from squidpy_b200._dist import world, shard_range, all_gather_rows, all_reduce_sum
    from squidpy_b200._rng import spawn_states
    from oracle import ref
    from tools import synth
    rank, ws = world()
    assert ws == 2
    # permutations sharded like nhood_enrichment does it; the oracle stands in for the per-rank GPU work
    g = synth.hex_graph(13, 17); n = g.shape[0]
    lab = np.random.default_rng(0).integers(0, 4, n).astype(np.uint32)
    P = 11
    lo, hi = shard_range(P, rank, ws)
    local = ref.nhood_perm_counts(g.indptr, g.indices, lab, 4, spawn_states(5, P, lo, hi))
    full = all_gather_rows(local, P)
    exp = ref.nhood_perm_counts(g.indptr, g.indices, lab, 4, spawn_states(5, P))
    assert full.shape == exp.shape and (full == exp).all(), "gathered permutation counts differ"
    # multi-GPU z-score statistics without gathering the counts: exact integer sums + variance accumulation chained through the
    # ranks; numpy stands in for the two device kernels (sqb_nhood_permute_sums / _var_chain, same operation order)
    from squidpy_b200._dist import sequential_stats
    def var_step(mean, acc):
        x = local.astype(np.float64)
        for p in range(x.shape[0]):
            d = x[p] - mean
            acc = acc + d * d
        return acc
    lo, hi = shard_range(P, rank, ws)
    mean, std = sequential_stats(local.sum(axis=0, dtype=np.int64), var_step, P)
    ef = exp.astype(np.float64)
    assert mean.tobytes() == ef.mean(axis=0).tobytes(), "chained mean differs from numpy's"
    assert std.tobytes() == ef.std(axis=0).tobytes(), "chained std differs from numpy's"
    # float64 feature scores (spatial_autocorr sharding)
    sc = np.arange(7, dtype=np.float64) * 1.5
    lo, hi = shard_range(7, rank, ws)
    assert (all_gather_rows(sc[lo:hi].copy(), 7) == sc).all()
    # int64 partial pair counts (co_occurrence / ripley tile sharding)
    part = np.full((3, 3, 5), rank + 1, dtype=np.int64)
    assert (all_reduce_sum(part) == 3).all()
    # seed=None: every rank must end up with the SAME entropy (rank 0 draws, broadcast), an explicit seed passes through
    from squidpy_b200._dist import shared_seed
    s = shared_seed(None)
    both = all_gather_rows(np.array([[s >> 64, s & ((1 << 64) - 1)]], dtype=np.uint64), 2)
    assert s > (1 << 64) and (both[0] == both[1]).all(), "ranks drew different entropy for seed=None"
    assert shared_seed(7) == 7
    a = np.random.default_rng(np.random.SeedSequence(s).spawn(3)[2]).random(4)
    assert (all_gather_rows(a[None, :].copy(), 2)[0] == all_gather_rows(a[None, :].copy(), 2)[1]).all()
    dist.destroy_process_group()
    print("RANK_OK", rank)
    """
)


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, SQB_ROOT=ROOT, PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", str(script)]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert out.stdout.count("RANK_OK") == 2, out.stdout


def test_single_process_passthrough():
    from squidpy_b200._dist import all_gather_rows, all_reduce_sum, world

    assert world() == (0, 1)
    a = np.arange(6).reshape(3, 2)
    assert all_gather_rows(a, 3) is a and all_reduce_sum(a) is a
