#!/bin/bash
# round 2, call K (2 GPUs): multi-GPU parity with the host-streaming cases
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_multi.py -q -k two > gpurun_out/r2k_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log; grep -E "passed|failed|FAILED|rc=|Error|mismatch" gpurun_out/r2k_pytest.log | tail -20
