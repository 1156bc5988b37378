"""Count-kernel timing at configs[1] (1M spots, hex graph, symmetric pair list) and configs[4] (300k cells, directed kNN graph):
per-kernel-class CUDA-event times for every `count_un` (row records per warp pass) and for the CSR-row kernel (count_sym=0)."""
import sys

import numpy as np

sys.path.insert(0, ".")
import squidpy_b200 as sq  # noqa: E402
from squidpy_b200.gr import NhoodPlan  # noqa: E402
from tools import synth  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
ctx = sq.default_context(0)
graphs = [("hex 1M sym", synth.hex_graph(1000, 1000), 30)]
rng = np.random.default_rng(0)
graphs.append(("knn 300k directed", synth.knn_graph(rng.random((300_000, 2)), 6), 12))
for name, g, C in graphs:
    base = synth.categorical_labels(g.shape[0], C, seed=0).cat.codes.to_numpy().astype(np.uint32)
    plan = NhoodPlan(g.indptr, g.indices, C, ctx)
    plan.set_base(base)
    plan.upload_philox(0, 0, P)
    ref = None
    for opt in (("count_un", 1), ("count_un", 2), ("count_un", 3), ("count_un", 4), ("count_sym", 0)):
        plan.set_option(*opt)
        plan.run_async()
        ctx.sync()
        ctx.profile(True)
        ctx.profile_reset()
        plan.run_async()
        ctx.sync()
        ms = ctx.profile_get("count")[0]
        ctx.profile(False)
        got = plan.download()
        if ref is None:
            ref = got
        print(name, opt, "count %.3f ms" % ms, "equal" if np.array_equal(ref, got) else "DIFFERENT", flush=True)
    plan.close()
