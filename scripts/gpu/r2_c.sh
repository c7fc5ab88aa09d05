#!/bin/bash
# round 2, call C (2 GPUs): multi-GPU parity (1x2, 2x1 incl. TRANS, sub-matrix, RSRC/CSRC != 0), bench plumbing on 1x2, BASELINE C3 at 1x2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
true
true
timeout 600 $TR bench.py --gpus 2 --size 32768 --steps 2 --warmup 1 > gpurun_out/r2c_bench_weak32k.json 2> gpurun_out/r2c_bench_weak32k.err
echo "weak32k rc=$?"; tail -c 600 gpurun_out/r2c_bench_weak32k.err
timeout 600 $TR bench.py --gpus 2 --config c3 --size 32768 --steps 2 --warmup 1 > gpurun_out/r2c_bench_c3_32k.json 2> gpurun_out/r2c_bench_c3_32k.err
echo "c3-32k rc=$?"; tail -c 400 gpurun_out/r2c_bench_c3_32k.err
timeout 600 $TR bench.py --gpus 2 --config c5 --size 8192 --steps 2 --warmup 1 > gpurun_out/r2c_bench_c5_8k.json 2> gpurun_out/r2c_bench_c5_8k.err
echo "c5-8k rc=$?"; tail -c 400 gpurun_out/r2c_bench_c5_8k.err
timeout 900 $TR bench.py --gpus 2 --config c3 --steps 2 --warmup 1 --no-e2e > gpurun_out/r2c_bench_c3_1x2.json 2> gpurun_out/r2c_bench_c3_1x2.err
echo "c3 rc=$?"; tail -c 400 gpurun_out/r2c_bench_c3_1x2.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["config"]["workload"], "value", round(d["value"], 2), "pct", round(d["config"]["pct_of_fp64_tensor_peak"], 1), "e2e", d["e2e"] and d["e2e"].get("value"), "pg", d.get("e2e_pageable") and d["e2e_pageable"].get("value"),
              "same", d["e2e"] and d["e2e"].get("bit_identical_to_device_resident"), "solve_ms", d["roofline_solve"]["solve_ms"], "sresid", d["config"]["sresid"], "pre", d["parity_preflight"]["ok"])
    except Exception as e:
        print(f, "unreadable", e)
PY
