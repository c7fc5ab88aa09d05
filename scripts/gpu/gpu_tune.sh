#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { env "$@" timeout 600 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('$*', round(d['value'],3), round(d['ms_per_step'],1), round(d['roofline']['achieved'],2), round(d['roofline']['share_of_step'],3))"; }
run SLB200_PANEL_GMAX=48 SLB200_GEMM_CHUNK=8
run SLB200_PANEL_GMAX=24 SLB200_GEMM_CHUNK=8
run SLB200_PANEL_GMAX=16 SLB200_GEMM_CHUNK=8
run SLB200_PANEL_GMAX=32 SLB200_GEMM_CHUNK=4
run SLB200_PANEL_GMAX=32 SLB200_GEMM_CHUNK=16
run SLB200_PANEL_GMAX=74 SLB200_GEMM_CHUNK=8
run SLB200_PANEL_GMAX=32 SLB200_GEMM_CHUNK=8 SLB200_LOOKAHEAD_MIN_US=1500
run SLB200_PANEL_GMAX=32 SLB200_GEMM_CHUNK=8 SLB200_LOOKAHEAD_MIN_US=8000
