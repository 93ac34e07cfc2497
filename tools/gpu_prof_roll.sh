#!/bin/bash
# ncu --set full capture of the row-streaming body-layer kernels at the cfg4 shape (8 images per launch):
# conv_roll_d_kernel (default) and conv_roll_kernel (--tc-variant 256).
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --workload cfg4 --batch 8 --e2e-iters 1 --no-cpu-baseline --no-hbm-probe"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"conv_roll_d_kernel" -s 20 -c 1 -o gpurun_out/prof_conv_roll_d -f $CMD > gpurun_out/ncu_full_d.log 2>&1
echo "rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"conv_roll_kernel" -s 20 -c 1 -o gpurun_out/prof_conv_roll -f $CMD --tc-variant 256 > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"; ls -la gpurun_out/*.ncu-rep 2>/dev/null
