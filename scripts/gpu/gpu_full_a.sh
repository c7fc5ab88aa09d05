#!/bin/bash
# Call A: full GPU parity suite (1 GPU), smoke, the default bench line (e2e + cpu_baseline) and the reference arm.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
for f in tests/test_gpu_kernels.py tests/test_gpu_lu.py tests/test_gpu_multi.py; do
  b=$(basename $f .py)
  timeout 600 python -m pytest $f -m gpu -q --timeout 240 --timeout-method=thread -p no:cacheprovider > gpurun_out/$b.log 2>&1
  echo "$f exit=$? $(tail -n 1 gpurun_out/$b.log)" | tee -a gpurun_out/summary.txt
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$? $(tail -n 1 gpurun_out/smoke.log)" | tee -a gpurun_out/summary.txt
s=$(date +%s)
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench rc=$? wall=$(( $(date +%s) - s ))s" | tee -a gpurun_out/summary.txt
s=$(date +%s)
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "reference rc=$? wall=$(( $(date +%s) - s ))s" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench_default.json gpurun_out/bench_reference.json | cut -c1-1800
tail -n 3 gpurun_out/bench_default.err
