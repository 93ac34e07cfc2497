#!/bin/bash
# ncu --set full capture of the 2-CTA body-layer kernel, the first and the last layer at the cfg4 shape (8 images per launch).
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --workload cfg4 --batch 8 --e2e-iters 1 --no-cpu-baseline --no-hbm-probe"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"conv_tc2_kernel|conv_first_tc_kernel|conv_tc_kernel" -s 40 -c 21 -o gpurun_out/prof_conv_layers -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"; ls -la gpurun_out/*.ncu-rep 2>/dev/null
