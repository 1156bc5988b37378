"""Small driver for ncu captures of the non-headline kernels at benchmark shapes (one target per run)."""
import sys, numpy as np
sys.path.insert(0, ".")
import squidpy_b200 as sq
from tools import synth

target = sys.argv[1]
ctx = sq.default_context(0)
if target == "moran":
    from sklearn.preprocessing import normalize
    from squidpy_b200.gr import AutocorrPlan
    g = synth.hex_graph(447, 447); normalize(g, norm="l1", axis=1, copy=False)
    x = synth.expression_csr(g.shape[0], 2048, density=0.1, coords=synth.hex_coords(447, 447), seed=1)
    plan = AutocorrPlan(g, ctx); plan.load(x, obs_major=True)
    for _ in range(2):
        s = plan.score("moran")
    print("moran ok", np.isfinite(s).sum())
elif target == "cooc":
    from squidpy_b200.gr import cooc_counts
    rng = np.random.default_rng(4); n = 200_000
    pts = (rng.random((n, 2)) * 2.0e4).astype(np.float32); labs = rng.integers(0, 20, n).astype(np.int32)
    thr = np.linspace(50.0, 1.4e4, 50, dtype=np.float32)[1:] ** 2
    c = cooc_counts(pts[:, 0], pts[:, 1], thr, labs, 20, ctx=ctx)
    print("cooc ok", int(c[:, :, -1].sum()))
elif target == "ripley":
    from squidpy_b200.gr import pair_counts
    pts = synth.thomas_points(150_000, seed=5); lab = synth.dirichlet_labels(150_000, 12, seed=5).cat.codes.to_numpy()
    sup = np.linspace(0, 7000, 50)
    c = pair_counts([pts[lab == k] for k in range(12)], sup, ctx=ctx)
    print("ripley ok", int(c[:, -1].sum()))
elif target == "shuffle":
    # nhood shuffle kernel alone at the headline shape: python tools/prof_targets.py shuffle P key=value ...
    from squidpy_b200._rng import spawn_states
    from squidpy_b200.gr import NhoodPlan
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 296
    g = synth.hex_graph(1000, 1000)
    base = synth.categorical_labels(g.shape[0], 30, seed=0).cat.codes.to_numpy().astype(np.uint32)
    plan = NhoodPlan(g.indptr, g.indices, 30, ctx)
    for kv in sys.argv[3:]:
        k, v = kv.split("=")
        plan.set_option(k, int(v))
    plan.set_base(base)
    plan.upload(spawn_states(0, P))
    plan.run_async(); ctx.sync()
    print("shuffle ok", int(plan.download().sum()))
