#!/bin/bash
# round 2, call D (1 GPU): parity suite, default bench with timeline, ncu launch list + full captures of swap / panel / update / solve kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_lu.py::test_full_size_properties_n65536 > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log; grep -E "passed|failed|FAILED|rc=" gpurun_out/r2d_pytest.log | tail -12
for m in laswp panel solve; do timeout 300 python scripts/ncu_driver.py $m > gpurun_out/r2d_drv_$m.log 2>&1; tail -1 gpurun_out/r2d_drv_$m.log; done
SLB200_LA_TRACE=1 timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
echo "bench rc=$?"; grep "la_trace: total" gpurun_out/r2d_bench.err | tail -3
NCU="ncu --clock-control none"
timeout 600 $NCU --set full --import-source on -k regex:swap_ -s 4 -c 3 -f -o gpurun_out/r2d_swap python scripts/ncu_driver.py laswp > gpurun_out/r2d_ncu_swap.log 2>&1; echo "ncu swap rc=$?"
timeout 600 $NCU --set full --import-source on -k regex:panel_leaf -s 127 -c 2 -f -o gpurun_out/r2d_leaf python scripts/ncu_driver.py panel > gpurun_out/r2d_ncu_leaf.log 2>&1; echo "ncu leaf rc=$?"
timeout 900 $NCU --set full --import-source on -k regex:"gemv_rows|diag_solve" -s 776 -c 6 -f -o gpurun_out/r2d_solve python scripts/ncu_driver.py solve > gpurun_out/r2d_ncu_solve.log 2>&1; echo "ncu solve rc=$?"
timeout 600 $NCU --set full --import-source on -k regex:dgemm_minus_packed -s 1 -c 1 -f -o gpurun_out/r2d_gemm python scripts/gemm_driver.py 32768 32768 512 1 > gpurun_out/r2d_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 900 $NCU --metrics gpu__time_duration.sum -c 60000 --csv --log-file gpurun_out/r2d_launches_n16384.csv python bench.py --size 16384 --steps 1 --warmup 0 --no-e2e --no-cpu --no-preflight > gpurun_out/r2d_ncu_list.log 2>&1; echo "list rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2d_bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["e2e"].get("bit_identical_to_device_resident"), "pageable", d["e2e_pageable"]["value"],
          "solve_ms", d["roofline_solve"]["solve_ms"], "solve frac", d["roofline_solve"]["frac"], "cpu", d["cpu_baseline"]["value"])
except Exception as e:
    print("bench unreadable", e)
PY
ls -la gpurun_out/*.ncu-rep 2>/dev/null
