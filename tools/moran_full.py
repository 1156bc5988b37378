"""configs[2] at full size (199 809 spots x G genes CSR f32 @10 %): load (H2D + device re-layout), run, download timings and
per-kernel-class CUDA-event times.  python tools/moran_full.py [G] [n_perms]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import squidpy_b200 as sq  # noqa: E402
from sklearn.preprocessing import normalize  # noqa: E402
from squidpy_b200.gr import AutocorrPlan  # noqa: E402
from tools import synth  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
NP = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ctx = sq.default_context(0)
g = synth.hex_graph(447, 447)
normalize(g, norm="l1", axis=1, copy=False)
t0 = time.perf_counter()
x = synth.expression_csr(g.shape[0], G, density=0.1, coords=synth.hex_coords(447, 447), seed=1)
print("synth %.1fs nnz=%d" % (time.perf_counter() - t0, x.nnz), flush=True)
plan = AutocorrPlan(g, ctx)
for rep in range(3):
    t0 = time.perf_counter()
    plan.load(x, obs_major=True)
    t1 = time.perf_counter()
    plan.run_async("moran")
    s = plan.download()
    t2 = time.perf_counter()
    print("rep %d: load %.3fs run+download %.4fs" % (rep, t1 - t0, t2 - t1), flush=True)
ctx.profile(True)
for mode in ("moran", "geary"):
    ctx.profile_reset()
    s1 = plan.score(mode)
    ctx.sync()
    print(mode, "main %.3f ms" % ctx.profile_get("autocorr_main")[0], "finite", int(np.isfinite(s1).sum()), flush=True)
ctx.profile_reset()
plan.load(x, obs_major=True)
ctx.sync()
print("load kernels (prep) %.3f ms over %d launches" % ctx.profile_get("autocorr_prep"), flush=True)
ctx.profile(False)
if NP:
    rng = np.random.default_rng(0)
    idx = np.stack([rng.permutation(g.shape[0]) for _ in range(NP)])
    t0 = time.perf_counter()
    sp = plan.score_perms("moran", idx)
    print("score_perms(%d) %.3fs" % (NP, time.perf_counter() - t0), "finite", int(np.isfinite(sp).sum()), flush=True)
