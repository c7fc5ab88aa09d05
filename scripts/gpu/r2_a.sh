#!/bin/bash
# round 2, call A (1 GPU): parity suite with the new interface tests, smoke, bench at a reduced and at the full size
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_gpu.txt 2>&1
nproc >> gpurun_out/r2a_gpu.txt; free -g >> gpurun_out/r2a_gpu.txt
which gfortran mpif90 mpirun >> gpurun_out/r2a_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_lu.py::test_full_size_properties_n65536 > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/r2a_smoke.log 2>&1; tail -2 gpurun_out/r2a_smoke.log
timeout 600 python bench.py --size 16384 --steps 2 --warmup 1 > gpurun_out/r2a_bench16k.json 2> gpurun_out/r2a_bench16k.err
echo "bench16k rc=$?"; tail -c 1500 gpurun_out/r2a_bench16k.err
SLB200_LA_TRACE=0 timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"; tail -c 800 gpurun_out/r2a_bench.err
python - <<'PY'
import json
for f in ("gpurun_out/r2a_bench16k.json", "gpurun_out/r2a_bench.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", d["value"], "e2e", d["e2e"], "pageable", d["e2e_pageable"], "solve", d["roofline_solve"]["solve_ms"], "frac", d["roofline"]["frac"], "pre", d["parity_preflight"]["ok"])
    except Exception as e:
        print(f, "unreadable", e)
PY
