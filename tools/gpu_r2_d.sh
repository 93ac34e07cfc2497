#!/bin/bash
set -u
mkdir -p gpurun_out
CMD="python bench.py --workload cfg1 --steps 3 --warmup 3 --e2e-iters 1 --no-cpu-baseline --no-hbm-probe"
ncu --set full --clock-control none --import-source on -k regex:"conv_chain_kernel" -s 4 -c 1 -o gpurun_out/prof_chain -f $CMD > gpurun_out/ncu_chain.log 2>&1
echo "rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"conv_tc2_kernel" -s 40 -c 1 -o gpurun_out/prof_tc2 -f $CMD --tc-variant 512 > gpurun_out/ncu_tc2.log 2>&1
echo "rc=$?"; ls -la gpurun_out/*.ncu-rep
