#!/bin/bash
# Runs the GPU parity tests file by file (a hang in one file cannot take the others down).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
free -g >> gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
for f in tests/test_gpu_kernels.py tests/test_gpu_lu.py; do
  b=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q -s --timeout 240 --timeout-method=thread -p no:cacheprovider > gpurun_out/$b.log 2>&1
  echo "$f exit=$?" | tee -a gpurun_out/summary.txt
  tail -5 gpurun_out/$b.log
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" | tee -a gpurun_out/summary.txt
grep -h "PEAKS\|passed\|failed" gpurun_out/*.log | head -20
