mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ac_sparse_kernel" -s 1 -c 1 -f -o gpurun_out/r02_prof_moran python tools/prof_targets.py moran > gpurun_out/ncu_moran.log 2>&1; echo "ncu moran rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"nhood_philox_labels|nhood_count_kernel" -s 2 -c 2 -f -o gpurun_out/r02_prof_philox python tools/philox_time.py 1000 > gpurun_out/ncu_philox.log 2>&1; echo "ncu philox rc=$?"
tail -3 gpurun_out/ncu_philox.log
timeout 300 python -m pytest tests/test_gpu_sepal.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
