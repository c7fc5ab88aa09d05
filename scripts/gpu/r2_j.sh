#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 300 python scripts/ncu_driver.py solve 2>&1 | tail -1
SLB200_SOLVE_CARVEOUT=0 timeout 300 python scripts/ncu_driver.py solve 2>&1 | tail -1
timeout 300 python scripts/ncu_driver.py solve 16384 512 2>&1 | tail -1
