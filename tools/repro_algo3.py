import sys, numpy as np
sys.path.insert(0, ".")
import squidpy_b200 as sq
from squidpy_b200._rng import spawn_states
from squidpy_b200.gr import NhoodPlan
from oracle import ref
import scipy.sparse as sp
n, P, nt, r = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
opts = dict(kv.split("=") for kv in sys.argv[5:])
g = sp.csr_matrix((np.ones(n - 1, np.float32), (np.arange(n - 1), np.arange(1, n))), shape=(n, n))
base = (np.arange(n) % 30).astype(np.uint32)
plan = NhoodPlan(g.indptr, g.indices, 30)
plan.set_option("shuffle_algo", 3); plan.set_option("shuffle_threads", nt); plan.set_option("shuffle_r", r)
for k, v in opts.items():
    plan.set_option(k, int(v))
plan.set_base(base)
st = spawn_states(5, P)
try:
    out = plan.permute(st)
    print("n", n, "P", P, nt, r, opts, "perm ok", (out.reshape(P,-1).sum(1) == g.nnz).all())
except Exception as e:
    print("n", n, "P", P, nt, r, opts, "FAILED", str(e)[-60:])
