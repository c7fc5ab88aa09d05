#!/bin/bash
# Call E: parity of the two-half pipeline + timeline + bench (1 GPU).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lu.py -m gpu -q -x -k "pipelined or lookahead or large_properties or example or ludat" --timeout 300 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -n 15 | tee gpurun_out/pipe_parity.log
SLB200_LA_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/trace_pipe.json 2> gpurun_out/trace_pipe.err
grep la_trace gpurun_out/trace_pipe.err | tail -n 42
run() {
  env "$@" timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu 2> gpurun_out/tune_pipe.err | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('$*', round(d['value'],3), round(d['ms_per_step'],1), round(d['roofline']['achieved'],2), round(d['roofline']['share_of_step'],3), d['config']['sresid'])"
}
{
run SLB200_LA_PIPELINE=1
run SLB200_LA_PIPELINE=0
run SLB200_SWAP_GRID=96
run SLB200_SWAP_GRID=0
run SLB200_GEMM_CHUNK=2
run SLB200_GEMM_CHUNK=8
} 2>&1 | tee gpurun_out/tune_pipe.txt
