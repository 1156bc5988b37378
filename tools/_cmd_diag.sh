mkdir -p gpurun_out
N=${1:-4}
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 tools/dist_diag.py > gpurun_out/dist_diag_${N}.txt 2> gpurun_out/dist_diag_${N}.err; echo "diag rc=$?"
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/dist_diag_${N}.txt | cut -c1-420; tail -n 4 gpurun_out/dist_diag_${N}.err | cut -c1-300
nvidia-smi topo -m 2>/dev/null | head -12
