#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lu.py -q -k "host_resident" > gpurun_out/r2e_pytest.log 2>&1; grep -E "passed|failed|FAILED" gpurun_out/r2e_pytest.log | tail
timeout 300 python scripts/e2e_debug.py 16384 512 > gpurun_out/r2e_dbg16k.log 2>&1; cat gpurun_out/r2e_dbg16k.log | tail -12
timeout 300 python scripts/e2e_debug.py 16384 512 e2e_slab_mb=0 > gpurun_out/r2e_dbg16k_slab0.log 2>&1; cat gpurun_out/r2e_dbg16k_slab0.log | tail -12
timeout 300 python scripts/e2e_debug.py 32768 512 > gpurun_out/r2e_dbg32k.log 2>&1; cat gpurun_out/r2e_dbg32k.log | tail -12
