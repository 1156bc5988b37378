mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ligrec.py tests/test_gpu_autocorr.py tests/test_gpu_graphs.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_b.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_b.log | cut -c1-300
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_ligrec.py tests/test_gpu_autocorr.py -m gpu -x -q -p no:cacheprovider -k "golden or perm_batch" > gpurun_out/san_b.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/san_b.log
timeout 600 python tools/moran_full.py 20000 100 > gpurun_out/moran_full.log 2>&1; tail -8 gpurun_out/moran_full.log
