#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/gpus.txt
timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "four" --timeout 900 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -30 | tee gpurun_out/multi4.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --size 32768 --steps 1 --warmup 1 --no-e2e --profile > gpurun_out/bench4_small.json 2> gpurun_out/bench4_small.err
echo "bench4 rc=$?"; cat gpurun_out/bench4_small.json; tail -5 gpurun_out/bench4_small.err
