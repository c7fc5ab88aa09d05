#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
./scripts/dmma_probe2 > gpurun_out/dmma_probe2.txt 2>&1
python scripts/gemm_driver.py 16384 16384 512 5 > gpurun_out/gemm_time.txt 2>&1
python scripts/gemm_driver.py 65024 8192 512 3 >> gpurun_out/gemm_time.txt 2>&1
python scripts/gemm_driver.py 8192 8192 256 5 1 >> gpurun_out/gemm_time.txt 2>&1
cat gpurun_out/gemm_time.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dgemm_minus -s 1 -c 1 -f -o gpurun_out/prof_gemm python scripts/gemm_driver.py 16384 16384 512 1 > gpurun_out/ncu_gemm.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_gemm.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_n8192.csv python bench.py --size 8192 --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/ncu_bench.log 2>&1
echo "ncu2 rc=$?"
cat gpurun_out/dmma_probe2.txt
