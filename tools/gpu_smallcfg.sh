#!/bin/bash
# refresh the small-workload bench lines (cfg1 / cfg2 / cfg3) with the CPU baseline
set -u
mkdir -p gpurun_out
for w in cfg1 cfg2 cfg3; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --e2e-iters 100 --no-hbm-probe > gpurun_out/BENCH_$w.json 2> gpurun_out/BENCH_$w.err; echo "$w rc=$?"; tail -c 200 gpurun_out/BENCH_$w.json
done
