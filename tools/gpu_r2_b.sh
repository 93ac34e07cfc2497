#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests/ -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
grep -E "^BASE_|^UNS_|FAILED" gpurun_out/pytest_gpu.log | tail -30
timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 --e2e-iters 10 --no-hbm-probe --no-cpu-baseline > gpurun_out/BENCH_cfg5.json 2> gpurun_out/BENCH_cfg5.err; echo "cfg5 rc=$?"; tail -2 gpurun_out/BENCH_cfg5.err; cat gpurun_out/BENCH_cfg5.json | head -c 3000
