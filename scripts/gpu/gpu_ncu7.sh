#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dgemm_minus_p8b -s 1 -c 1 -f -o gpurun_out/prof_gemm_v7 python scripts/gemm_driver.py 32768 32768 512 1 > gpurun_out/ncu_gemm_v7.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_gemm_v7.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file gpurun_out/launches_n16384.csv python bench.py --size 16384 --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/ncu_bench_16384.log 2>&1
echo "list rc=$?"
