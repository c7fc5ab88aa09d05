#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lu.py -m gpu -q -x -k "variants" --timeout 300 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -3
for v in 7 8; do
  echo "variant $v"; SLB200_GEMM_VARIANT=$v python scripts/gemm_driver.py 16384 16384 512 5; SLB200_GEMM_VARIANT=$v python scripts/gemm_driver.py 65024 4096 512 5; SLB200_GEMM_VARIANT=$v python scripts/gemm_driver.py 2048 2048 512 20; SLB200_GEMM_VARIANT=$v python scripts/gemm_driver.py 448 65536 64 20
done 2>&1 | tee gpurun_out/gemm_ab4.txt
SLB200_GEMM_VARIANT=8 timeout 900 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('v8 bench', round(d['value'],3), round(d['ms_per_step'],1), round(d['roofline']['achieved'],2), round(d['roofline']['share_of_step'],3), d['config']['sresid'])"
