mkdir -p gpurun_out
timeout 75 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29566 bench.py --gpus 4 --steps 5 --warmup 3 --skip-moran --skip-pairs --skip-cpu > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err; echo "rc=$?"
python tools/print_bench.py gpurun_out/bench_4gpu.json 2>&1 | head -4
