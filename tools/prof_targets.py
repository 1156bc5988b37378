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
elif target == "misc":
    # one launch each of the kNN search (1M points, k=6), the sepal diffusion (71x71 lattice, 64 genes) and the ligrec group means
    import pandas as pd
    from squidpy_b200.gr import knn_2d, ligrec_analysis, sepal_scores
    from squidpy_b200.gr._sepal import _compute_idxs
    d, i = knn_2d(synth.hex_coords(1000, 1000), 6, ctx=ctx)
    g = synth.hex_graph(71, 71); co = synth.hex_coords(71, 71)
    rng = np.random.default_rng(0)
    sat, sat_idx, unsat, unsat_idx = _compute_idxs(g, co, 6)
    sc = sepal_scores(rng.random((g.shape[0], 64)) * rng.integers(1, 6, 64), sat, sat_idx, unsat, unsat_idx, max_neighs=6, n_iter=2000, ctx=ctx)
    x = rng.poisson(0.8, (50000, 256)).astype(np.float64); cl = rng.integers(0, 12, 50000)
    df = pd.DataFrame(x, columns=list(range(256))); df["clusters"] = pd.Categorical(cl, categories=list(range(12)))
    inter = np.stack([rng.integers(0, 256, 500), rng.integers(0, 256, 500)], 1); cp = np.array([(a, b) for a in range(12) for b in range(12)])
    r = ligrec_analysis(df, inter, cp, threshold=0.05, n_perms=200, seed=1, ctx=ctx)
    print("misc ok", int(i.sum()), float(np.nanmean(sc)), float(np.nanmean(r.pvalues)))
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
