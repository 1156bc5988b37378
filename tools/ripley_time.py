import sys, time, numpy as np
sys.path.insert(0, ".")
import squidpy_b200 as sq
from squidpy_b200.gr import pair_counts
from tools import synth
from scipy.spatial import ConvexHull
ctx = sq.default_context(0)
pts = synth.thomas_points(300_000, seed=5)
lab = synth.dirichlet_labels(300_000, 12, seed=5).cat.codes.to_numpy()
area = ConvexHull(pts).volume
support = np.linspace(0, (area / 2) ** 0.5, 50)
groups = [pts[lab == c] for c in range(12)]
pair_counts([g[:2000] for g in groups], support, ctx=ctx)
for rep in range(4):
    ctx.profile(rep == 3); ctx.profile_reset()
    t0 = time.perf_counter(); c = pair_counts(groups, support, ctx=ctx); dt = time.perf_counter() - t0
    print("rep", rep, "seconds %.4f" % dt, "pairs kernel ms", ctx.profile_get("pairs") if rep == 3 else "")
