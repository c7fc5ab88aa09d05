#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lu.py tests/test_gpu_kernels.py -m gpu -q -x --timeout 300 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/trip_test.log
for v in 3 6 7; do
  echo "variant $v"; SLB200_GEMM_VARIANT=$v python scripts/gemm_driver.py 16384 16384 512 5; SLB200_GEMM_VARIANT=$v python scripts/gemm_driver.py 65024 4096 512 5
done 2>&1 | tee gpurun_out/gemm_ab3.txt
for v in 3 7; do
SLB200_GEMM_VARIANT=$v timeout 900 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
echo "v=$v rc=$?"; cat gpurun_out/bench_v$v.json | python -c "import json,sys; d=json.load(sys.stdin); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['share_of_step'], d['config']['sresid'])"; tail -3 gpurun_out/bench_v$v.err
done
