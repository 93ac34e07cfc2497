#!/bin/bash
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --workload cfg4 --batch 8 --e2e-iters 1 --no-cpu-baseline --no-hbm-probe"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"conv_first_tc_kernel" -s 2 -c 1 -o gpurun_out/prof_first_tc $CMD > gpurun_out/ncu_first.log 2>&1
echo "rc=$?"
PDS_TC_VARIANT=64 $CMD > gpurun_out/plain2.log 2>&1 && \
PDS_TC_VARIANT=64 ncu --set full --clock-control none --import-source on -k regex:"conv_first_kernel" -s 2 -c 1 -o gpurun_out/prof_first_simt $CMD > gpurun_out/ncu_first2.log 2>&1
echo "rc=$?"; ls -la gpurun_out/*.ncu-rep
