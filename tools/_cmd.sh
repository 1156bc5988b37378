mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_philox.py tests/test_gpu_autocorr.py tests/test_autocorr_known_answers.py -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_a.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_a.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_philox.py tests/test_gpu_autocorr.py -m gpu -x -q -p no:cacheprovider -k "spec or library or shards or weight_formats or sparse_layouts" > gpurun_out/san_a.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/san_a.log
timeout 300 python tools/philox_time.py > gpurun_out/philox_time.log 2>&1; tail -3 gpurun_out/philox_time.log
timeout 300 python tools/moran_time.py > gpurun_out/moran_time.log 2>&1; tail -2 gpurun_out/moran_time.log
timeout 600 python tools/moran_full.py 20000 20 > gpurun_out/moran_full.log 2>&1; tail -12 gpurun_out/moran_full.log
