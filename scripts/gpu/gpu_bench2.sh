#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 1 --profile > gpurun_out/bench2.json 2> gpurun_out/bench2.err
echo "bench2 rc=$?"; cat gpurun_out/bench2.json; tail -5 gpurun_out/bench2.err
