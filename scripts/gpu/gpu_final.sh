#!/bin/bash
# Final 1-GPU pass: parity suite (minus the slow subprocess A/B cases), smoke, default bench line, reference arm.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 300 python -m pytest tests -m gpu -q -k "not variants and not packed" --timeout 240 --timeout-method=thread -p no:cacheprovider > gpurun_out/final_tests.log 2>&1
echo "tests exit=$? $(tail -n 1 gpurun_out/final_tests.log)" | tee -a gpurun_out/summary.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$? $(tail -n 1 gpurun_out/smoke.log)" | tee -a gpurun_out/summary.txt
s=$(date +%s)
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench rc=$? wall=$(( $(date +%s) - s ))s" | tee -a gpurun_out/summary.txt
timeout 200 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "reference rc=$?" | tee -a gpurun_out/summary.txt
cut -c1-2600 gpurun_out/bench_default.json; tail -n 3 gpurun_out/bench_default.err
