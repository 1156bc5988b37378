mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_autocorr.py tests/test_autocorr_known_answers.py -m gpu -x -q -p no:cacheprovider -k "not million" > gpurun_out/san_autocorr.log 2>&1; echo "memcheck autocorr rc=$?"
tail -15 gpurun_out/san_autocorr.log
timeout 600 python -m pytest tests/test_gpu_autocorr.py tests/test_autocorr_known_answers.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_autocorr.log 2>&1; echo "pytest autocorr rc=$?"
tail -30 gpurun_out/pytest_autocorr.log
timeout 300 python tools/moran_time.py > gpurun_out/moran_time.log 2>&1; cat gpurun_out/moran_time.log | tail -5
