#!/bin/bash
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --workload cfg4 --batch 8 --e2e-iters 1 --no-cpu-baseline --no-hbm-probe"
ncu --set full --clock-control none --import-source on -k regex:"blur_rt_kernel" -s 10 -c 2 -o gpurun_out/prof_blur -f $CMD > gpurun_out/ncu_full_blur.log 2>&1
echo "rc=$?"; ls -la gpurun_out/prof_blur.ncu-rep 2>/dev/null
