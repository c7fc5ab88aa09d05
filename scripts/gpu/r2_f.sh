#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_lu.py::test_full_size_properties_n65536 > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log; grep -E "passed|failed|FAILED|rc=" gpurun_out/r2f_pytest.log | tail -12
timeout 300 python scripts/e2e_debug.py 32768 512 > gpurun_out/r2f_dbg32k.log 2>&1; tail -4 gpurun_out/r2f_dbg32k.log
for m in panel solve; do timeout 300 python scripts/ncu_driver.py $m > gpurun_out/r2f_drv_$m.log 2>&1; tail -1 gpurun_out/r2f_drv_$m.log; done
timeout 300 python scripts/ncu_driver.py solve 16384 512 2>&1 | tail -1
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2f_bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["e2e"].get("bit_identical_to_device_resident"), "pageable", d["e2e_pageable"]["value"], d["e2e_pageable"].get("bit_identical_to_device_resident"),
          "solve_ms", d["roofline_solve"]["solve_ms"], "solve frac", d["roofline_solve"]["frac"], "cpu", d["cpu_baseline"]["value"])
except Exception as e:
    print("bench unreadable", e)
PY
