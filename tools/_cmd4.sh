mkdir -p gpurun_out
N=${1:-4}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench$N rc=$?"
python tools/print_bench.py gpurun_out/bench_${N}gpu.json; nproc; tail -n 3 gpurun_out/bench_${N}gpu.err | cut -c1-200
