#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/gpus.txt
for v in 2 3 4 5; do
  echo "variant $v"; SLB200_GEMM_VARIANT=$v python scripts/gemm_driver.py 16384 16384 512 5
done 2>&1 | tee gpurun_out/gemm_ab2.txt
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --timeout 600 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -30 | tee gpurun_out/multi.log
