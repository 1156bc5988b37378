import sys, numpy as np
sys.path.insert(0, ".")
from squidpy_b200._rng import spawn_states
from squidpy_b200.gr import NhoodPlan
import scipy.sparse as sp
n, P, nt, r = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
lo = int(sys.argv[5]) if len(sys.argv) > 5 else 0
algo = int(sys.argv[6]) if len(sys.argv) > 6 else 3
g = sp.csr_matrix((np.ones(n - 1, np.float32), (np.arange(n - 1), np.arange(1, n))), shape=(n, n))
base = (np.arange(n) % 30).astype(np.uint32)
plan = NhoodPlan(g.indptr, g.indices, 30)
plan.set_option("shuffle_algo", algo); plan.set_option("shuffle_threads", nt); plan.set_option("shuffle_r", r)
plan.set_base(base)
st = spawn_states(5, P)
for p in range(lo, P):
    try:
        plan.permute(st[p:p+1])
    except Exception as e:
        print("FAILED at perm", p, str(e)[-50:], flush=True)
        break
else:
    print("all single perms ok", lo, P, "algo", algo, nt, r)
