#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_lu.py::test_full_size_properties_n65536 > gpurun_out/r2g_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log; grep -E "passed|failed|FAILED|rc=" gpurun_out/r2g_pytest.log | tail -12
timeout 300 python scripts/ncu_driver.py solve 2>&1 | tail -1
timeout 300 python scripts/ncu_driver.py solve 16384 512 2>&1 | tail -1
SLB200_SOLVE_GRAPH=0 timeout 300 python scripts/ncu_driver.py solve 16384 512 2>&1 | tail -1
for v in 16384 16384,4 8192; do timeout 120 python scripts/gemm_driver.py ${v%,*} ${v%,*} 256 3 1 2>&1 | tail -1; done
SLB200_ZGEMM_PACKED=0 timeout 120 python scripts/gemm_driver.py 16384 16384 256 3 1 2>&1 | tail -1
timeout 600 python bench.py --config c5 --size 24576 --steps 2 --warmup 1 --no-pageable > gpurun_out/r2g_bench_c5_24k.json 2> gpurun_out/r2g_bench_c5_24k.err; echo "c5 rc=$?"
SLB200_ZGEMM_PACKED=0 timeout 600 python bench.py --config c5 --size 24576 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2g_bench_c5_24k_old.json 2> gpurun_out/r2g_bench_c5_24k_old.err; echo "c5 old rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2g_bench_c5_24k.json", "gpurun_out/r2g_bench_c5_24k_old.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", d["value"], "pct", d["config"]["pct_of_fp64_tensor_peak"], "frac", d["roofline"]["frac"], "e2e", d["e2e"] and d["e2e"]["value"], "sresid", d["config"]["sresid"])
    except Exception as e:
        print(f, "unreadable", e)
PY
