mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_graphs.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_graphs.log 2>&1; echo "pytest graphs rc=$?"
tail -40 gpurun_out/pytest_graphs.log | cut -c1-400
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_graphs.py -m gpu -x -q -p no:cacheprovider -k "golden or known or degenerate" > gpurun_out/san_graphs.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/san_graphs.log
python - <<'PY'
import time, numpy as np, sys
sys.path.insert(0,'.')
from tools import synth
from squidpy_b200.gr import GridBuilder, KNNBuilder, knn_2d
co = synth.hex_coords(1000, 1000)
for rep in range(3):
    t0=time.perf_counter(); d,i,m = knn_2d(co, 6, median=True); t1=time.perf_counter()
    adj,dst = GridBuilder(n_neighs=6).build(co); t2=time.perf_counter()
    print("knn_2d 1M k=6: %.3fs; GridBuilder.build: %.3fs; nnz %d" % (t1-t0, t2-t1, adj.nnz), flush=True)
rng=np.random.default_rng(0); co2=rng.random((1000000,2))*1e4
t0=time.perf_counter(); adj,dst = KNNBuilder(n_neighs=6).build(co2); print("KNNBuilder 1M random: %.3fs" % (time.perf_counter()-t0))
from sklearn.neighbors import NearestNeighbors
t0=time.perf_counter(); NearestNeighbors(n_neighbors=6).fit(co2[:200000]).kneighbors(); print("sklearn kneighbors 200k: %.3fs" % (time.perf_counter()-t0))
PY
