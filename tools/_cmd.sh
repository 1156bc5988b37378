mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_gpu.log
timeout 300 python tools/philox_time.py > gpurun_out/philox_time.log 2>&1; tail -3 gpurun_out/philox_time.log
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 9000 gpurun_out/bench.json; tail -n 15 gpurun_out/bench.err
