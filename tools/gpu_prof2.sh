#!/bin/bash
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --workload cfg4 --batch 8 --e2e-iters 1 --no-cpu-baseline --no-hbm-probe"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"conv_first_tc_kernel|conv_tc2_kernel" -s 10 -c 3 -o gpurun_out/prof_first_tc2 $CMD > gpurun_out/ncu_first.log 2>&1
echo "rc=$?"; ls -la gpurun_out/prof_first_tc2.ncu-rep
