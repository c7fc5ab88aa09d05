#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file gpurun_out/launches_full.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/ncu_bench_full.log 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/ncu_bench_full.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dgemm_minus_persistent -s 3 -c 1 -f -o gpurun_out/prof_gemm_v3 python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/ncu_gemm_v3.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/ncu_gemm_v3.log | cut -c1-300
timeout 600 ncu --set full --clock-control none -k regex:panel_leaf -s 40 -c 1 -f -o gpurun_out/prof_leaf python bench.py --size 16384 --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/ncu_leaf.log 2>&1
echo "leaf rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:swap_pack -s 20 -c 1 -f -o gpurun_out/prof_swap python bench.py --size 32768 --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/ncu_swap.log 2>&1
echo "swap rc=$?"
ls -la gpurun_out/*.ncu-rep
