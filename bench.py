#!/usr/bin/env python
"""bench.py — throughput of the squidpy spatial-statistics hot path on B200.

Headline (BASELINE.json `metric`, quoted on configs[1]): permutations/s of `nhood_enrichment` on 1 000 000 spots,
30 clusters, k=6 hexagonal neighbour graph (nnz = 5 992 002), n_perms = 1000 per GPU, exact numpy-RNG replay.
A "step" = one pass of the permutation test (1000 permutations: fill + shuffle + transpose + count kernels) with the
graph, base labels and generator states already resident in HBM.  `e2e` = the same metric through the public API
(`squidpy_b200.gr.nhood_enrichment(adata, ...)`) with host buffers: H2D of CSR/labels/states, kernels, D2H of the
per-permutation counts and the float64 z-score on the host, all inside the timed region.
Extras on the same JSON line: Moran's I genes/s on configs[2] (200k spots x 20k genes CSR, sharded by genes under
torchrun), and with --all co_occurrence (configs[3]) / Ripley L (configs[4]) pair rates.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--all] [--skip-moran]
N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG2 = dict(rows=1000, cols=1000, n_cls=30, n_perms=1000, seed=0)
CFG3 = dict(rows=447, cols=447, n_genes=20000, density=0.1)
METRIC = "nhood_enrichment permutations/s (1M spots, 30 clusters, k=6, n_perms=1000/GPU, exact numpy-RNG replay)"
WORKLOAD = "configs[1]: 1M-spot hex grid (nnz=5992002), 30 clusters, k=6, nhood_enrichment n_perms=1000 per GPU"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel (the Fisher-Yates apply kernel) per launch, from
    the committed `ncu --set full` capture of this same workload (profiles/r01_prof_nhood_metrics.csv); None if absent."""
    import csv

    path = os.path.join(ROOT, "profiles", "r01_prof_nhood_metrics.csv")
    try:
        rows = {r[0]: r for r in csv.reader(open(path))}
        names = rows["metric"][2:]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        dur = [float(v) for v in rows["gpu__time_duration.sum"][2:]]
        col = 2 + max(range(len(names)), key=lambda i: dur[i])  # the longest launch of the capture
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(rows[k][col]) * scale[rows[k][1]]
        return {"bytes_per_launch": tot, "kernel": names[col - 2], "source": "profiles/r01_prof_nhood_metrics.csv (ncu --set full, P=1000)"}
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for k, nm in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def _dist_setup(n_gpus: int):
    import torch

    rank, ws, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if ws > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return rank, ws, local


def _barrier_sync(ws):
    import torch

    if ws > 1:
        import torch.distributed as dist

        dist.barrier()
    torch.cuda.synchronize()


def _max_over_ranks(v: float, ws: int) -> float:
    if ws == 1:
        return v
    import torch
    import torch.distributed as dist

    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---------------------------------------------------------------------------------------------------------------
def cpu_reference_nhood(g, base, n_cls, seed, budget_s=20.0):
    """The reference's CPU algorithm (oracle C port: (N,C)-scratch two-pass count + exact numpy shuffle) on the host
    cores: all threads over permutations == joblib n_jobs=-1 semantics.  Bounded sample."""
    from oracle import ref
    from squidpy_b200._rng import spawn_states

    import oracle

    cores = len(os.sched_getaffinity(0))
    ref.nhood_perm_counts(g.indptr, g.indices, base, n_cls, spawn_states(seed, 2), n_threads=cores)  # warm-up
    t0 = time.perf_counter()
    ref.nhood_perm_counts(g.indptr, g.indices, base, n_cls, spawn_states(seed, 1), n_threads=1)
    t1 = time.perf_counter() - t0
    p_s = int(max(cores, min(12 * cores, budget_s / max(t1, 1e-3) * cores * 0.6)))
    t0 = time.perf_counter()
    ref.nhood_perm_counts(g.indptr, g.indices, base, n_cls, spawn_states(seed, p_s), n_threads=cores)
    dt = time.perf_counter() - t0
    return {"value": p_s / dt, "unit": "permutations/s", "cores": cores, "kind": "port",
            "sample": f"{p_s} permutations of the same 1M-spot workload on {cores} OpenMP threads ({dt:.1f}s); serial: {1.0 / t1:.2f} perm/s",
            "serial_value": 1.0 / t1}


def run_reference(args, rank, ws):
    """--impl reference: the reference's own CPU path (oracle port; the Python/numba reference cannot travel to the
    GPU box) on the host cores, same config/metric; each step = a bounded sample of the workload."""
    if rank != 0:
        return
    from oracle import ref
    from squidpy_b200._rng import spawn_states
    from tools import synth

    g = synth.hex_graph(CFG2["rows"], CFG2["cols"])
    base = np.random.default_rng(0).integers(0, CFG2["n_cls"], g.shape[0]).astype(np.uint32)
    cores = len(os.sched_getaffinity(0))
    p_s = max(cores, 2 * cores)
    for _ in range(max(args.warmup, 1)):
        ref.nhood_perm_counts(g.indptr, g.indices, base, CFG2["n_cls"], spawn_states(0, cores), n_threads=cores)
    t0 = time.perf_counter()
    for k in range(args.steps):
        ref.nhood_perm_counts(g.indptr, g.indices, base, CFG2["n_cls"], spawn_states(k, p_s), n_threads=cores)
    dt = time.perf_counter() - t0
    val = args.steps * p_s / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "permutations/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic", "config": {"workload": WORKLOAD, "sample_perms_per_step": p_s},
            "cpu_baseline": {"value": val, "unit": "permutations/s", "cores": cores, "kind": "port",
                             "sample": f"{p_s} permutations/step x {args.steps} steps, {cores} OpenMP threads over permutations (joblib n_jobs=-1 semantics)"},
            "e2e": {"value": val, "unit": "permutations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
def bench_moran(ctx, rank, ws, steps, warmup, flush):
    """configs[2]: 199 809 spots x 20 000 genes CSR float32 @10 %, mode='moran'; genes sharded over ranks (strong)."""
    import torch
    from sklearn.preprocessing import normalize

    from squidpy_b200._dist import shard_range
    from squidpy_b200.gr import AutocorrPlan
    from tools import synth

    g = synth.hex_graph(CFG3["rows"], CFG3["cols"])
    normalize(g, norm="l1", axis=1, copy=False)
    n = g.shape[0]
    lo, hi = shard_range(CFG3["n_genes"], rank, ws)
    t0 = time.perf_counter()
    x = synth.expression_csr(n, hi - lo, density=CFG3["density"], coords=synth.hex_coords(CFG3["rows"], CFG3["cols"]), seed=100 + rank)
    t_gen = time.perf_counter() - t0
    plan = AutocorrPlan(g, ctx)
    _barrier_sync(ws)
    t0 = time.perf_counter()
    plan.load(x, obs_major=True)
    plan.run_async("moran")
    score = plan.download()
    t_e2e = _max_over_ranks(time.perf_counter() - t0, ws)
    for _ in range(warmup):
        plan.run_async("moran")
    _barrier_sync(ws)
    tot = 0.0
    l0 = ctx.launches
    for _ in range(steps):
        flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.run_async("moran")
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    _barrier_sync(ws)
    launches = ctx.launches - l0
    ms = _max_over_ranks(tot / steps, ws)
    ctx.profile(True)
    ctx.profile_reset()
    plan.run_async("moran")
    ctx.sync()
    kms = {k: ctx.profile_get(k)[0] for k in ("autocorr_prep", "autocorr_main", "autocorr_final")}
    ctx.profile(False)
    nnz_x = x.nnz
    algo_bytes = 8 * nnz_x + 8 * (hi - lo + 1) + 8 * g.nnz + 4 * (n + 1) + 8 * (hi - lo)
    peak, _ = _peaks()
    out = {"metric": "Moran's I genes/s (199 809 spots x 20 000 genes CSR f32 @10%, genes sharded over GPUs)", "value": CFG3["n_genes"] / (ms / 1e3),
           "unit": "genes/s", "ms_per_step": ms, "scaling": "strong", "genes_per_rank": hi - lo, "nnz_x_rank0": int(nnz_x),
           "e2e": {"value": CFG3["n_genes"] / t_e2e, "unit": "genes/s", "h2d_bytes_per_step": int(x.data.nbytes + x.indices.nbytes + 8 * (n + 1)),
                   "d2h_bytes_per_step": int(8 * (hi - lo)), "seconds": t_e2e, "note": "load (H2D + device CSR transposition) + run + download, pageable scipy buffers"},
           "roofline": {"bound": "hbm", "achieved": algo_bytes / (ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s", "frac": algo_bytes / (ms / 1e3) / 1e9 / peak,
                        "algorithmic_bytes_per_step": int(algo_bytes)},
           "finite_scores": int(np.isfinite(score).sum()), "max_I": float(np.nanmax(score)), "gpu_launches": int(launches), "synth_seconds": t_gen,
           "kernel_ms": kms}
    if rank == 0 and ws == 1:  # CPU baseline at N = 1 only
        try:
            from oracle import ref

            cores = len(os.sched_getaffinity(0))
            ns = 2 * cores
            t0 = time.perf_counter()
            exp = ref.morans_i(g, x[:, :ns].T.tocsr(), n_threads=cores)
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": ns / dt, "unit": "genes/s", "cores": cores, "kind": "port",
                                   "sample": f"{ns} genes of the same matrix, scanpy-style restatement (oracle C, OpenMP over genes), {dt:.1f}s"}
            out["parity_max_abs_err_sample"] = float(np.nanmax(np.abs(score[:ns] - exp)))
        except Exception as e:  # pragma: no cover
            out["cpu_baseline"] = {"error": repr(e)}
    plan.close()
    return out


def bench_cooc(ctx, flush):
    """configs[3]: 500 000 uniform points, 20 clusters, 50 radii (L = 49)."""
    import torch

    from squidpy_b200.gr import cooc_counts
    from squidpy_b200.gr._ppatterns import _find_min_max

    rng = np.random.default_rng(4)
    n = 500_000
    pts = (rng.random((n, 2)) * 2.0e4).astype(np.float32)
    labs = rng.integers(0, 20, n).astype(np.int32)
    tmin, tmax = _find_min_max(pts)
    iv = np.linspace(tmin, tmax, 50, dtype=np.float32)
    thr = iv[1:] ** 2
    cooc_counts(pts[:20000, 0], pts[:20000, 1], thr, labs[:20000], 20, ctx=ctx)  # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    c = cooc_counts(pts[:, 0], pts[:, 1], thr, labs, 20, ctx=ctx)
    dt = time.perf_counter() - t0
    pairs = float(n) * (n - 1)
    return {"metric": "co_occurrence ordered pairs/s (500k points, 20 clusters, 49 radii)", "value": pairs / dt, "unit": "ordered pairs/s", "seconds": dt,
            "note": "through the C ABI with host buffers (H2D/D2H inside); symmetry used: n(n-1)/2 pair evaluations", "within_max_radius_frac": float(c[:, :, -1].sum() / pairs),
            "bound": "FP32/INT issue rate, not HBM (each point re-used ~1000x from shared memory)"}


def bench_ripley(ctx):
    """configs[4]: 300 000 MERFISH-shaped cells, 12 clusters, Ripley L pair counting (float64), 50 radii."""
    from scipy.spatial import ConvexHull

    from squidpy_b200.gr import pair_counts
    from tools import synth

    pts = synth.thomas_points(300_000, seed=5)
    lab = synth.dirichlet_labels(300_000, 12, seed=5).cat.codes.to_numpy()
    area = ConvexHull(pts).volume
    support = np.linspace(0, (area / 2) ** 0.5, 50)
    groups = [pts[lab == c] for c in range(12)]
    import gc

    pair_counts([g[:2000] for g in groups], support, ctx=ctx)
    gc.collect()  # handles of the previous benchmarks (GB-sized device buffers) are not freed inside the timed call
    ctx.sync()
    times = []
    for _ in range(3):
        t0 = time.perf_counter()
        pair_counts(groups, support, ctx=ctx)
        times.append(time.perf_counter() - t0)
    dt = min(times)
    pairs = float(sum(len(g) ** 2 for g in groups))
    return {"metric": "Ripley L ordered pairs/s (300k cells, 12 clusters, float64, 50 radii)", "value": pairs / dt, "unit": "ordered pairs/s", "seconds": dt,
            "seconds_all": times, "note": "best of 3 calls through the C ABI with host buffers", "largest_cluster": int(max(len(g) for g in groups))}


def bench_nhood_variants(ctx, g, base, n_cls, P, seed):
    """configs[1] variants of SURVEY 8(d): the same lattice with randomly permuted node order (worst-case gather locality
    for the count kernel) and with half of the mirrored entries dropped (a directed graph: no symmetric shortcut)."""
    import scipy.sparse as sp

    from squidpy_b200._rng import spawn_states
    from squidpy_b200.gr import NhoodPlan

    out = {}
    rng = np.random.default_rng(1)
    n = g.shape[0]
    perm = rng.permutation(n)
    g_perm = g[perm][:, perm].tocsr()
    g_perm.sort_indices()
    coo = g.tocoo()
    keep = (coo.row < coo.col) | (rng.random(coo.nnz) < 0.5)  # drops ~half of the j < i mirrors
    g_dir = sp.csr_matrix((coo.data[keep], (coo.row[keep], coo.col[keep])), shape=g.shape)
    states = spawn_states(seed, P)
    for name, gg, lab in (("permuted_node_order", g_perm, base[perm]), ("directed_graph", g_dir, base)):
        plan = NhoodPlan(gg.indptr, gg.indices, n_cls, ctx)
        plan.set_base(lab)
        plan.upload(states)
        plan.run_async()
        ctx.sync()
        ctx.profile(True)
        ctx.profile_reset()
        plan.run_async()
        ctx.sync()
        kms = {k: ctx.profile_get(k)[0] for k in ("fill", "misc", "shuffle", "transpose", "count")}
        ctx.profile(False)
        kms["jgen"] = kms.pop("misc")
        counts = plan.download()
        assert (counts.reshape(P, -1).sum(axis=1, dtype=np.int64) == gg.nnz).all()
        tot = sum(kms.values())
        out[name] = {"nnz": int(gg.nnz), "kernel_ms": kms, "permutations_per_s": P / (tot / 1e3),
                     "note": "sum of per-kernel CUDA-event times of one step (launches synchronised)"}
        plan.close()
    return out


# ---------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--all", action="store_true", help="also run co_occurrence (configs[3]) and Ripley L (configs[4])")
    ap.add_argument("--skip-moran", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--perms", type=int, default=CFG2["n_perms"])
    ap.add_argument("--shuffle-threads", type=int, default=0)
    ap.add_argument("--shuffle-algo", type=int, default=-1)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        rank, ws = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
        run_reference(args, rank, ws)
        return

    import torch

    import squidpy_b200 as sq
    from squidpy_b200._rng import spawn_states
    from squidpy_b200.gr import NhoodPlan
    from tools import synth

    rank, ws, local = _dist_setup(args.gpus)
    # the library launches on a torch-owned stream made current here, so torch.cuda.Event timing sees its kernels
    # (torch's default stream has handle 0, which the C ABI reads as "create your own stream")
    bench_stream = torch.cuda.Stream()
    torch.cuda.set_stream(bench_stream)
    ctx = sq.Context(local, bench_stream.cuda_stream)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def flush():
        flush_buf.add_(1)  # 256 MiB read+write > 126 MB L2

    g = synth.hex_graph(CFG2["rows"], CFG2["cols"])
    n, n_cls, P = g.shape[0], CFG2["n_cls"], args.perms
    labels = synth.categorical_labels(n, n_cls, seed=0)
    base = labels.cat.codes.to_numpy().astype(np.uint32)
    plan = NhoodPlan(g.indptr, g.indices, n_cls, ctx)
    if args.shuffle_threads:
        plan.set_option("shuffle_threads", args.shuffle_threads)
    plan.set_option("shuffle_algo", args.shuffle_algo)
    plan.set_base(base)
    states = spawn_states(CFG2["seed"], P * ws, rank * P, (rank + 1) * P)  # this rank's generators of the N*P-permutation job
    plan.upload(states)
    for _ in range(args.warmup):
        plan.run_async()
    _barrier_sync(ws)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launches
    tot = 0.0
    for _ in range(args.steps):
        flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.run_async()
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    _barrier_sync(ws)
    launches = ctx.launches - l0
    clocks = sampler.stop() if rank == 0 else None
    ms_step = _max_over_ranks(tot / args.steps, ws)
    value = ws * P / (ms_step / 1e3)
    counts = plan.download()
    assert (counts.reshape(P, -1).sum(axis=1, dtype=np.int64) == g.nnz).all(), "count checksum failed"

    # ---- roofline pass: per-kernel-class CUDA-event times of ONE step (launches synchronised, not part of `value`)
    ctx.profile(True)
    ctx.profile_reset()
    plan.run_async()
    ctx.sync()
    kms = {k: ctx.profile_get(k)[0] for k in ("fill", "misc", "shuffle", "transpose", "count")}
    kms["jgen"] = kms.pop("misc")  # swap-target generation (PCG64 replay + rejection sampling), accounted as class "misc"
    ctx.profile(False)
    peak, peak_src = _peaks()
    bpp = plan.bytes_per_perm
    dom = max(kms, key=kms.get)
    step_gbs = bpp * P / (ms_step / 1e3) / 1e9
    roofline = {"bound": "hbm", "achieved": step_gbs, "peak": peak, "unit": "GB/s", "frac": step_gbs / peak, "traffic": _ncu_traffic(),
                "kernel": "nhood permutation step = fill + jgen + shuffle (apply) + transpose + count (one launch each per 1000 permutations); dominant: nhood_apply_list_kernel",
                "algorithmic_bytes_per_perm": int(bpp), "bytes_formula": "4*nnz + 4*(N+1) + 8*N + 4*C^2 (SURVEY.md 8d, reference dtypes)", "peak_source": peak_src,
                "kernel_ms": kms, "dominant_kernel": dom, "dominant_share": kms[dom] / max(sum(kms.values()), 1e-9)}

    # ---- e2e through the public API with host buffers (H2D + kernels + D2H + host z-score in the timed region)
    ad = synth.make_adata(np.zeros((n, 2)), g, labels)
    e2e_steps = max(1, min(args.steps, 3))
    sq.gr.nhood_enrichment(ad, "cluster", n_perms=P * ws, seed=CFG2["seed"], copy=True, device=local)
    _barrier_sync(ws)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = sq.gr.nhood_enrichment(ad, "cluster", n_perms=P * ws, seed=CFG2["seed"], copy=True, device=local)
    _barrier_sync(ws)
    t_e2e = _max_over_ranks((time.perf_counter() - t0) / e2e_steps, ws)
    h2d = int(g.indptr.nbytes + g.indices.nbytes + 2 * base.nbytes + states.nbytes)
    d2h = int(P * n_cls * n_cls * 4 + n_cls * n_cls * 4)
    e2e = {"value": ws * P / t_e2e, "unit": "permutations/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "seconds_per_call": t_e2e,
           "note": "sq.gr.nhood_enrichment(adata, n_perms, seed, copy=True): CSR + labels + PCG64 states H2D (pageable scipy/numpy buffers), kernels, counts D2H, float64 z-score on host"}
    assert np.isfinite(res.zscore).all()

    extras = {}
    if not args.skip_moran:
        try:
            extras["moran"] = bench_moran(ctx, rank, ws, max(1, min(args.steps, 3)), 1, flush)
        except Exception as e:  # pragma: no cover
            extras["moran"] = {"error": repr(e)}
    if args.all and rank == 0 and ws == 1:
        for name, fn in (("co_occurrence", lambda: bench_cooc(ctx, flush)), ("ripley_L", lambda: bench_ripley(ctx)),
                         ("nhood_variants", lambda: bench_nhood_variants(ctx, g, base, n_cls, P, CFG2["seed"]))):
            try:
                extras[name] = fn()
            except Exception as e:  # pragma: no cover
                extras[name] = {"error": repr(e)}

    if rank == 0:
        cpu = None
        if not args.skip_cpu and ws == 1:  # the CPU baseline is timed on rank 0 at N = 1 only
            try:
                cpu = cpu_reference_nhood(g, base, n_cls, CFG2["seed"])
            except Exception as e:  # pragma: no cover
                cpu = {"error": repr(e)}
        line = {"metric": METRIC, "value": value, "unit": "permutations/s", "n_gpus": ws, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8 labels / u32 counts", "data": "synthetic",
                "config": {"workload": WORKLOAD, "n_perms_per_gpu": P, "rng": "numpy PCG64 exact replay", "shuffle_algo": args.shuffle_algo,
                           "l2": "flushed between steps (256 MiB read+write); working set per step 2 GB > L2", "timing": "CUDA events per step on the launch stream, max over ranks"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "extras": extras}
        print(json.dumps(line), flush=True)
    plan.close()
    if ws > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
