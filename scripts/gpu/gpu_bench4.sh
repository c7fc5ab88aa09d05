#!/bin/bash
# 4-GPU (2x2 grid): light parity (BASELINE config 1 + a pipelined split case), then the default-size bench line.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 120 python scripts/run_mp.py 4 '[{"P":2,"Q":2,"m":2000,"n":2000,"nb":64,"nrhs":1},{"P":2,"Q":2,"m":4096,"n":4096,"nb":128,"nrhs":1,"dev":true,"split":256}]' 80 2>&1 | tail -n 8 | cut -c1-300 | tee gpurun_out/mp4.log
s=$(date +%s)
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 1 --warmup 1 --no-e2e > gpurun_out/bench4_pipe.json 2> gpurun_out/bench4_pipe.err
echo "bench4 rc=$? wall=$(( $(date +%s) - s ))s"; cut -c1-1300 gpurun_out/bench4_pipe.json; tail -n 3 gpurun_out/bench4_pipe.err | cut -c1-300
