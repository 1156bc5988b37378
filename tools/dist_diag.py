"""Per-rank anatomy of the multi-GPU nhood step (torchrun): kernels / collective / statistics times and launch skew of every
rank, plus bare NCCL collectives of the same sizes.  usage: torchrun --nproc-per-node N tools/dist_diag.py [fast]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import squidpy_b200 as sq  # noqa: E402
from squidpy_b200._rng import spawn_states  # noqa: E402
from squidpy_b200.gr import NhoodPlan  # noqa: E402
from tools import synth  # noqa: E402

rank, ws, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = sq.Context(local, stream.cuda_stream)
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
g = synth.hex_graph(1000, 1000)
base = synth.categorical_labels(g.shape[0], 30, seed=0).cat.codes.to_numpy().astype(np.uint32)
P, cc = 1000, 900
plan = NhoodPlan(g.indptr, g.indices, 30, ctx)
plan.set_base(base)
fast = len(sys.argv) > 1 and sys.argv[1] == "fast"
if fast:
    plan.upload_philox(0, rank * P, P)
else:
    plan.upload(spawn_states(0, P * ws, rank * P, (rank + 1) * P))
local_t = torch.zeros((P, cc), dtype=torch.int32, device="cuda")
full = torch.empty((ws * P, cc), dtype=torch.int32, device="cuda")
stat = torch.empty((2, cc), dtype=torch.float64, device="cuda")
import ctypes as C  # noqa: E402


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def step(collective=True):
    flush_buf.add_(1)
    t0 = time.perf_counter()
    e0 = ev()
    plan.run_async()
    e1 = ev()
    plan.counts_dev(local_t.data_ptr())
    if collective:
        dist.all_gather_into_tensor(full, local_t)
    e2 = ev()
    plan._lib.sqb_nhood_stats_rows_dev(ctx.handle, C.c_void_p(full.data_ptr()), ws * P, 30, C.c_void_p(stat[0].data_ptr()), C.c_void_p(stat[1].data_ptr()))
    e3 = ev()
    t1 = time.perf_counter()
    stat.cpu()
    t2 = time.perf_counter()
    return e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3), (t1 - t0) * 1e3, (t2 - t0) * 1e3


for _ in range(3):
    step()
dist.barrier()
torch.cuda.synchronize()
for mode in (True, False):
    rows = [step(mode) for _ in range(6)]
    dist.barrier()
    torch.cuda.synchronize()
    print(f"rank {rank} collective={mode} kernels/collective/stats/host-launch/host-total ms:", " | ".join("%.2f %.2f %.2f %.2f %.2f" % r for r in rows), flush=True)
# bare collectives
tiny = torch.zeros(900, dtype=torch.int64, device="cuda")
for name, fn in (("all_reduce 7KB", lambda: dist.all_reduce(tiny)), ("all_gather 3.6MB/rank", lambda: dist.all_gather_into_tensor(full, local_t)), ("barrier", lambda: dist.barrier())):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        a = ev()
        w0 = time.perf_counter()
        fn()
        b = ev()
        b.synchronize()
        ts.append((a.elapsed_time(b), (time.perf_counter() - w0) * 1e3))
    print(f"rank {rank} bare {name}: device/host ms", " ".join("%.3f/%.3f" % t for t in ts), flush=True)
if rank == 0:
    print("nproc", len(os.sched_getaffinity(0)), "NCCL", torch.cuda.nccl.version(), flush=True)
plan.close()
dist.destroy_process_group()
