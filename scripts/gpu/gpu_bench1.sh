#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
./scripts/dmma_probe > gpurun_out/dmma_probe.txt 2>&1
for n in 8192 16384 32768; do
  timeout 600 python bench.py --size $n --steps 2 --warmup 1 --profile --no-e2e --no-cpu > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  echo "n=$n rc=$?"; tail -c 1500 gpurun_out/bench_n$n.json
done
timeout 900 python bench.py --steps 2 --warmup 1 --profile > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
echo "full rc=$?"; cat gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
