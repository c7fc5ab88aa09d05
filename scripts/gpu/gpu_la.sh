#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lu.py tests/test_gpu_kernels.py -m gpu -q -x --timeout 300 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/la_test.log
for la in 0 1; do
SLB200_LOOKAHEAD=$la timeout 900 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_la$la.json 2> gpurun_out/bench_la$la.err
echo "la=$la rc=$?"; cat gpurun_out/bench_la$la.json | python -c "import json,sys; d=json.load(sys.stdin); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['share_of_step'], d['config']['sresid'])"; tail -3 gpurun_out/bench_la$la.err
done
SLB200_LOOKAHEAD=1 timeout 600 python bench.py --size 16384 --steps 2 --warmup 1 --no-e2e --no-cpu | python -c "import json,sys; d=json.load(sys.stdin); print('n16384 la1', d['value'], d['ms_per_step'])"
SLB200_LOOKAHEAD=0 timeout 600 python bench.py --size 16384 --steps 2 --warmup 1 --no-e2e --no-cpu | python -c "import json,sys; d=json.load(sys.stdin); print('n16384 la0', d['value'], d['ms_per_step'])"
