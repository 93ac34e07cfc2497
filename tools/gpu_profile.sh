#!/bin/bash
# Driver-style checks + profiles: full GPU test suite, default bench (+ reference arm), ncu launch list,
# one --set full capture of the body-layer kernel.
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/BENCH_ref.json 2> gpurun_out/BENCH_ref.err; tail -c 400 gpurun_out/BENCH_ref.json
echo "== bench default"; timeout 900 python bench.py > gpurun_out/BENCH_default.json 2> gpurun_out/BENCH_default.err; echo "rc=$?"; tail -c 1500 gpurun_out/BENCH_default.json; tail -3 gpurun_out/BENCH_default.err
CMD="python bench.py --steps 2 --warmup 3 --workload cfg4 --batch 8 --e2e-iters 1 --no-cpu-baseline --no-hbm-probe"
echo "== ncu launch list"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "rc=$?"; wc -l gpurun_out/launches.csv
echo "== ncu full"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 20 -c 3 -o gpurun_out/prof_conv_tc $CMD > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"; ls -la gpurun_out/*.ncu-rep 2>/dev/null
