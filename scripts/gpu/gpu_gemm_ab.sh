#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in 1 2; do
  echo "variant $v"
  SLB200_GEMM_VARIANT=$v python scripts/gemm_driver.py 16384 16384 512 5
  SLB200_GEMM_VARIANT=$v python scripts/gemm_driver.py 65024 16384 512 3
  SLB200_GEMM_VARIANT=$v python scripts/gemm_driver.py 4096 4096 512 10
  SLB200_GEMM_VARIANT=$v python scripts/gemm_driver.py 448 65536 64 10
done 2>&1 | tee gpurun_out/gemm_ab.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm or microbench" -s 2>&1 | tail -5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dgemm_minus_persistent -s 1 -c 1 -f -o gpurun_out/prof_gemm_v2 python scripts/gemm_driver.py 16384 16384 512 1 > gpurun_out/ncu_gemm_v2.log 2>&1
timeout 900 python bench.py --steps 2 --warmup 1 --profile --no-e2e --no-cpu > gpurun_out/bench_full_v2.json 2> gpurun_out/bench_full_v2.err
echo "full rc=$?"; cat gpurun_out/bench_full_v2.json | python -c "import json,sys; d=json.load(sys.stdin); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['peak'], d.get('phase_profile_us'))"
