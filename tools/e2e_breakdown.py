"""Where the end-to-end time of sq.gr.nhood_enrichment goes at configs[1] (host buffers in, z-scores out): wall-clock per stage."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import squidpy_b200 as sq  # noqa: E402
from squidpy_b200._rng import spawn_states  # noqa: E402
from squidpy_b200.gr import NhoodPlan  # noqa: E402
from squidpy_b200.gr._utils import as_csr  # noqa: E402
from tools import synth  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
ctx = sq.default_context(0)
g = synth.hex_graph(1000, 1000)
base = synth.categorical_labels(g.shape[0], 30, seed=0).cat.codes.to_numpy().astype(np.uint32)
for rep in range(3):
    t = [time.perf_counter()]
    adj = as_csr(g)
    t.append(time.perf_counter())
    plan = NhoodPlan(adj.indptr, adj.indices, 30, ctx)
    ctx.sync()
    t.append(time.perf_counter())
    cnt = plan.count(base)
    t.append(time.perf_counter())
    plan.set_base(base)
    ctx.sync()
    t.append(time.perf_counter())
    st = spawn_states(0, P, 0, P)
    t.append(time.perf_counter())
    plan.upload(st)
    ctx.sync()
    t.append(time.perf_counter())
    plan.run_async()
    ctx.sync()
    t.append(time.perf_counter())
    m, s = plan.stats()
    t.append(time.perf_counter())
    plan.close()
    t.append(time.perf_counter())
    names = ["as_csr", "create(H2D CSR, checks, records)", "observed count", "set_base", "spawn_states(host)", "upload states", "kernels", "stats+download", "close"]
    print("rep", rep, " ".join("%s=%.2fms" % (n, (b - a) * 1e3) for n, a, b in zip(names, t[:-1], t[1:])), "total=%.2fms" % ((t[-1] - t[0]) * 1e3), flush=True)
