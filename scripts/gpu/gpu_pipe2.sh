#!/bin/bash
# Call F: pipeline v2 (g0 kept clear of prep/left swaps, capped grids): parity + trace + tuning.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lu.py tests/test_gpu_kernels.py -m gpu -q -x -k "not variants and not packed" --timeout 300 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -n 8 | tee gpurun_out/pipe2_parity.log
SLB200_LA_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/trace_pipe2.json 2> gpurun_out/trace_pipe2.err
grep la_trace gpurun_out/trace_pipe2.err | tail -n 24
run() {
  env "$@" timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu 2> gpurun_out/tune_pipe2.err | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('$*', round(d['value'],3), round(d['ms_per_step'],1), round(d['roofline']['achieved'],2), round(d['roofline']['share_of_step'],3), d['config']['sresid'])"
}
{
run SLB200_SWAP_GRID=96
run SLB200_SWAP_GRID=48
run SLB200_SWAP_GRID=192
run SLB200_LA_PIPELINE=0
run SLB200_PANEL_GMAX=24
run SLB200_LA_SPLIT_MIN=16384
} 2>&1 | tee gpurun_out/tune_pipe2.txt
