#!/bin/bash
# Call B: parity + A/B timing of the packed-operand update kernel (variant 9) against v7.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lu.py -m gpu -q -x -k "packed" --timeout 120 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -n 15 | tee gpurun_out/v9_parity.log
{
echo "variant 7"; SLB200_GEMM_VARIANT=7 timeout 120 python scripts/gemm_driver.py 16384 16384 512 5
for lag in 0 3000 6000 12000 24000; do
  echo "variant 9 lag=$lag epi=0"; SLB200_GEMM_LAG=$lag timeout 120 python scripts/gemm_driver.py 16384 16384 512 5
done
echo "variant 9 lag=6000 epi=1"; SLB200_GEMM_EPI=1 timeout 120 python scripts/gemm_driver.py 16384 16384 512 5
echo "variant 9 lag=0 epi=1"; SLB200_GEMM_LAG=0 SLB200_GEMM_EPI=1 timeout 120 python scripts/gemm_driver.py 16384 16384 512 5
echo "variant 9 default, other shapes"
timeout 120 python scripts/gemm_driver.py 65024 4096 512 5
timeout 120 python scripts/gemm_driver.py 32768 512 512 10
timeout 120 python scripts/gemm_driver.py 4096 4096 512 10
echo "variant 7, other shapes"
SLB200_GEMM_VARIANT=7 timeout 120 python scripts/gemm_driver.py 32768 512 512 10
SLB200_GEMM_VARIANT=7 timeout 120 python scripts/gemm_driver.py 4096 4096 512 10
} 2>&1 | tee gpurun_out/gemm_ab9.txt
timeout 600 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_v9.json 2> gpurun_out/bench_v9.err
echo "bench rc=$?"; python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_v9.json') if l.startswith('{')][0]); print('v9 bench', round(d['value'],3), round(d['ms_per_step'],1), round(d['roofline']['achieved'],2), round(d['roofline']['share_of_step'],3), d['config']['sresid'])"
tail -n 3 gpurun_out/bench_v9.err
