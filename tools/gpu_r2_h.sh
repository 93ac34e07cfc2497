#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== full"; timeout 2000 python -m pytest tests/ -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.log
grep -E "^BASE_|^unet|FAILED|sparse part" gpurun_out/pytest_gpu.log | tail -30
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench default"; timeout 900 python bench.py > gpurun_out/BENCH_default.json 2> gpurun_out/BENCH_default.err; echo "rc=$?"; tail -c 1500 gpurun_out/BENCH_default.json; tail -3 gpurun_out/BENCH_default.err
