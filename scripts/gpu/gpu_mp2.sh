#!/bin/bash
# 2-GPU check of the pipelined P x Q path: light parity (1x2, 2x1) + the default 2-GPU bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 python scripts/run_mp.py 2 '[{"P":1,"Q":2,"m":1000,"n":1000,"nb":64,"nrhs":1,"dev":true},{"P":1,"Q":2,"m":3072,"n":3072,"nb":128,"nrhs":1,"dev":true,"split":256},{"P":2,"Q":1,"m":3072,"n":3072,"nb":128,"nrhs":1,"dev":true,"split":256},{"P":2,"Q":1,"m":777,"n":513,"nb":100,"nrhs":0},{"P":1,"Q":2,"m":120,"n":120,"nb":16,"nrhs":2,"z":true}]' 2>&1 | tail -n 12 | tee gpurun_out/mp2.log
s=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 1 --warmup 1 --no-e2e > gpurun_out/bench2_pipe.json 2> gpurun_out/bench2_pipe.err
echo "bench2 rc=$? wall=$(( $(date +%s) - s ))s"; cut -c1-1400 gpurun_out/bench2_pipe.json; tail -n 4 gpurun_out/bench2_pipe.err
