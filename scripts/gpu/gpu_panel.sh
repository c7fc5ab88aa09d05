#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "panel" -s --timeout 200 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/panel_test.log
timeout 600 python -m pytest tests/test_gpu_lu.py -m gpu -q -x --timeout 200 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -5 | tee -a gpurun_out/panel_test.log
timeout 900 python bench.py --steps 2 --warmup 1 --profile --no-e2e --no-cpu > gpurun_out/bench_full_v3.json 2> gpurun_out/bench_full_v3.err
echo "full rc=$?"; cat gpurun_out/bench_full_v3.json | python -c "import json,sys; d=json.load(sys.stdin); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['peak'], d.get('phase_profile_us'))"
