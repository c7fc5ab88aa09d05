#!/bin/bash
# 4-GPU (2x2 grid) default-size bench line (N=131072, 32 GiB of A per GPU).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/gpus.txt
s=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 1 --warmup 1 --no-e2e > gpurun_out/bench4.json 2> gpurun_out/bench4.err
echo "bench4 rc=$? wall=$(( $(date +%s) - s ))s"; cut -c1-1500 gpurun_out/bench4.json; tail -n 5 gpurun_out/bench4.err
