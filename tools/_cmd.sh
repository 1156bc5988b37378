mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sepal.py tests/test_gpu_ligrec.py tests/test_gpu_graphs.py -m gpu -q -p no:cacheprovider --durations=5 > gpurun_out/pytest_c.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_c.log | cut -c1-300
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_sepal.py -m gpu -x -q -p no:cacheprovider -k "golden and square" > gpurun_out/san_c.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/san_c.log
