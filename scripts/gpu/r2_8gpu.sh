#!/bin/bash
# round 2, the 8-GPU call: 2x4-grid parity (test_eight_gpus), BASELINE configs C4 (PDGETRF N=262144), C3 (PDGESV N=131072) and C5 (PZGETRF
# N=65536 NB=256) on the 2x4 grid, each with the parity pre-flight and the per-step pipeline trace.  Own evidence runs: 1 warm-up step only.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2_8gpu_gpus.txt; nproc >> gpurun_out/r2_8gpu_gpus.txt; free -g | head -2 >> gpurun_out/r2_8gpu_gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711"
timeout -k 10 600 python -m pytest tests/test_gpu_multi.py -q -k eight -s > gpurun_out/r2_parity_8gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_parity_8gpu.log; grep -E "passed|failed|FAILED|rc=" gpurun_out/r2_parity_8gpu.log | tail -5
SLB200_LA_TRACE=1 timeout -k 10 540 $TR bench.py --gpus 8 --config c4 --steps 1 --warmup 1 --no-e2e > gpurun_out/r2_bench_c4_2x4.json 2> gpurun_out/r2_bench_c4_2x4.err
echo "c4 rc=$?"
SLB200_LA_TRACE=1 timeout -k 10 300 $TR bench.py --gpus 8 --config c3 --steps 2 --warmup 1 --no-e2e > gpurun_out/r2_bench_c3_2x4.json 2> gpurun_out/r2_bench_c3_2x4.err
echo "c3 rc=$?"
timeout -k 10 300 $TR bench.py --gpus 8 --config c5 --steps 2 --warmup 1 --no-e2e > gpurun_out/r2_bench_c5_2x4.json 2> gpurun_out/r2_bench_c5_2x4.err
echo "c5 rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_bench_c*_2x4.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["config"]["workload"], "| value", round(d["value"], 2), "pct", round(d["config"]["pct_of_fp64_tensor_peak"], 1), "ms", round(d["ms_per_step"], 1), "frac", d["roofline"]["frac"],
              "share", d["roofline"]["share_of_step"], "solve_ms", d["roofline_solve"]["solve_ms"], "sresid", d["config"]["sresid"], "pre", d["parity_preflight"]["ok"], d["clocks"])
    except Exception as e:
        print(f, "unreadable", e)
PY
grep -h "la_trace\[0,0\]: total\|la_trace\[1,3\]: total" gpurun_out/r2_bench_c*_2x4.err | tail -8
