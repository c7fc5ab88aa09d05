#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "getrs or pdgesv or lu_dat or square or trans or submatrix or example" > gpurun_out/r2i_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log; grep -E "passed|failed|FAILED|rc=" gpurun_out/r2i_pytest.log | tail -12
timeout 300 python scripts/ncu_driver.py solve 2>&1 | tail -1
timeout 300 python scripts/ncu_driver.py solve 32768 512 2>&1 | tail -1
timeout 300 python scripts/ncu_driver.py solve 16384 512 2>&1 | tail -1
SLB200_SOLVE_GRAPH=0 timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,launch__grid_size -k regex:"diag_solve|gemv_rows|gemv_bulk|inv32" -s 0 -c 13 --csv --log-file gpurun_out/r2i_solve_kernels.csv python scripts/ncu_driver.py solve 32768 512 > gpurun_out/r2i.log 2>&1
grep -o '"[^"]*\(diag_solve\|gemv_rows\|gemv_bulk\|inv32\)[^"]\{0,30\}.*gpu__time_duration.sum","ns","[0-9]*"' gpurun_out/r2i_solve_kernels.csv | sed 's/"void slb::<unnamed>:://; s/(.*gpu__time/ gpu__time/' | head -14
