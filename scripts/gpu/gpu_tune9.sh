#!/bin/bash
# Call C: look-ahead timeline with the packed kernel, chunk / gmax tuning, ncu capture of the packed kernel.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() {  # env assignments as args
  env "$@" timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu 2> gpurun_out/tune9.err | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('$*', round(d['value'],3), round(d['ms_per_step'],1), round(d['roofline']['achieved'],2), round(d['roofline']['share_of_step'],3))"
}
SLB200_LA_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/trace9.json 2> gpurun_out/trace9.err
grep la_trace gpurun_out/trace9.err | tail -n 40
{
run SLB200_GEMM_LAG=0
run SLB200_GEMM_LAG=0 SLB200_GEMM_CHUNK=8
run SLB200_GEMM_LAG=0 SLB200_GEMM_CHUNK=2
run SLB200_GEMM_LAG=0 SLB200_PANEL_GMAX=16
run SLB200_GEMM_LAG=0 SLB200_PANEL_GMAX=48 SLB200_GEMM_CHUNK=8
run SLB200_GEMM_LAG=0 SLB200_LOOKAHEAD_MIN_US=1000
} 2>&1 | tee gpurun_out/tune9.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dgemm_minus_packed -s 1 -c 1 -f -o gpurun_out/prof_gemm_v9 python scripts/gemm_driver.py 32768 32768 512 1 > gpurun_out/ncu_gemm_v9.log 2>&1
echo "ncu rc=$?"; tail -n 2 gpurun_out/ncu_gemm_v9.log
