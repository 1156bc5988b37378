"""nhood step at configs[1] (1M spots, 30 clusters, P permutations) in exact and fast RNG mode: per-kernel-class CUDA-event times."""
import sys

import numpy as np

sys.path.insert(0, ".")
import squidpy_b200 as sq  # noqa: E402
from squidpy_b200._rng import spawn_states  # noqa: E402
from squidpy_b200.gr import NhoodPlan  # noqa: E402
from tools import synth  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
ctx = sq.default_context(0)
g = synth.hex_graph(1000, 1000)
base = synth.categorical_labels(g.shape[0], 30, seed=0).cat.codes.to_numpy().astype(np.uint32)
plan = NhoodPlan(g.indptr, g.indices, 30, ctx)
plan.set_base(base)
for mode in ("philox", "numpy"):
    if mode == "philox":
        plan.upload_philox(0, 0, P)
    else:
        plan.upload(spawn_states(0, P))
    plan.run_async()
    ctx.sync()
    ctx.profile(True)
    ctx.profile_reset()
    plan.run_async()
    ctx.sync()
    kms = {k: round(ctx.profile_get(k)[0], 3) for k in ("fill", "misc", "shuffle", "transpose", "count")}
    ctx.profile(False)
    m, s = plan.stats()
    print(mode, P, kms, "total %.3f ms" % sum(kms.values()), "mean[0,0]=%.2f std[0,0]=%.3f" % (m[0, 0], s[0, 0]), flush=True)
plan.close()
