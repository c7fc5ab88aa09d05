#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "two" --timeout 600 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/multi2b.log
for la in 1 0; do
SLB200_LOOKAHEAD=$la timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 1 --no-e2e > gpurun_out/bench2_la$la.json 2> gpurun_out/bench2_la$la.err
echo "bench2 la=$la rc=$?"; python -c "import json,sys; d=json.load(open('gpurun_out/bench2_la$la.json')); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['share_of_step'], d['config']['sresid'])"; tail -3 gpurun_out/bench2_la$la.err | cut -c1-300
done
