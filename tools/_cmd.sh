mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nhood.py tests/test_gpu_autocorr.py tests/test_gpu_philox.py tests/test_gpu_pairs.py -m gpu -x -q -p no:cacheprovider -k "count or autocorr or sparse or philox or api or unused or buffered" > gpurun_out/pytest_d.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_d.log | cut -c1-250
for g in 4 8; do SQB_AC_GRP=$g timeout 300 python tools/moran_time.py 2>&1 | tail -1; done
SQB_AC_GRP=4 timeout 600 python tools/moran_full.py 20000 0 2>&1 | grep -E "main|rep 2"
SQB_AC_GRP=8 timeout 600 python tools/moran_full.py 20000 0 2>&1 | grep -E "main|rep 2"
timeout 300 python tools/philox_time.py 2>&1 | tail -2
timeout 300 python tools/tune_nhood.py 1000 2>&1 | grep api_breakdown | cut -c1-600
