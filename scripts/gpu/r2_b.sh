#!/bin/bash
# round 2, call B (1 GPU): the whole 1-GPU parity suite (no -x)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/r2b_pytest.log | tail -30
