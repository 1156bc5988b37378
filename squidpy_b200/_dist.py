"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL on GPUs, gloo on CPU for tests).

The hot path shards embarrassingly — permutations for ``nhood_enrichment``, features for ``spatial_autocorr``, pair
tiles for ``co_occurrence``/``ripley`` — so there is exactly ONE collective per call, on the final small tensor
(SURVEY.md section 8e).  Chunking follows the reference's own rule for splitting work over workers:
contiguous chunks of ``ceil(n / n_workers)`` (``src/squidpy/_utils.py:225-231``).
"""

from __future__ import annotations

import numpy as np


def world() -> tuple[int, int]:
    """(rank, world_size) of the initialised default process group, (0, 1) otherwise."""
    try:
        import torch.distributed as dist
    except Exception:  # pragma: no cover
        return 0, 1
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n: int, rank: int, world_size: int) -> tuple[int, int]:
    step = -(-n // world_size) if n > 0 else 0
    lo = min(rank * step, n)
    return lo, min(lo + step, n)


def _device_for_backend():
    import torch
    import torch.distributed as dist

    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def all_gather_rows(local: np.ndarray, n_total: int) -> np.ndarray:
    """Concatenate per-rank row blocks (block ``r`` = ``shard_range(n_total, r, world)``) on every rank."""
    rank, ws = world()
    if ws == 1:
        return local
    import torch
    import torch.distributed as dist

    step = -(-n_total // ws)
    row_shape = local.shape[1:]
    pad = np.zeros((step,) + row_shape, dtype=local.dtype)
    pad[: local.shape[0]] = local
    dev = _device_for_backend()
    # NCCL has no uint32: move bytes
    t = torch.from_numpy(pad.view(np.uint8).reshape(-1)).to(dev)
    out = torch.empty(ws * t.numel(), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(out, t)
    full = out.cpu().numpy().view(local.dtype).reshape((ws * step,) + row_shape)
    return np.ascontiguousarray(full[:n_total])


def all_reduce_sum(local: np.ndarray) -> np.ndarray:
    """Element-wise sum over ranks (int64 / float64 payloads)."""
    rank, ws = world()
    if ws == 1:
        return local
    import torch
    import torch.distributed as dist

    dev = _device_for_backend()
    t = torch.from_numpy(np.ascontiguousarray(local)).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()
