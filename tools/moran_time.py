"""Moran main-kernel timing at a reduced shape (447x447 lattice, 2048 genes): per-kernel-class CUDA-event times of one
score() call.  Used to compare kernel variants quickly: python tools/moran_time.py"""
import sys

import numpy as np

sys.path.insert(0, ".")
import squidpy_b200 as sq  # noqa: E402
from sklearn.preprocessing import normalize  # noqa: E402
from squidpy_b200.gr import AutocorrPlan  # noqa: E402
from tools import synth  # noqa: E402

ctx = sq.default_context(0)
g = synth.hex_graph(447, 447)
normalize(g, norm="l1", axis=1, copy=False)
x = synth.expression_csr(g.shape[0], 2048, density=0.1, coords=synth.hex_coords(447, 447), seed=1)
plan = AutocorrPlan(g, ctx)
plan.load(x, obs_major=True)
s0 = plan.score("moran")
ctx.profile(True)
ctx.profile_reset()
s1 = plan.score("moran")
ctx.sync()
print("moran 2048 genes: prep %.3f main %.3f final %.3f ms; rerun identical %s; finite %d" % (
    ctx.profile_get("autocorr_prep")[0], ctx.profile_get("autocorr_main")[0], ctx.profile_get("autocorr_final")[0],
    bool(np.array_equal(s0, s1, equal_nan=True)), int(np.isfinite(s1).sum())))
