#!/bin/bash
# ncu captures of conv_roll_d_kernel at the cfg4 launch shape for the tc_variant values given as arguments (cycle counts are
# independent of the board's power state; used to compare epilogue store variants).
set -u
mkdir -p gpurun_out
for v in "$@"; do
  CMD="python bench.py --steps 2 --warmup 3 --workload cfg4 --batch 8 --e2e-iters 1 --no-cpu-baseline --no-hbm-probe --tc-variant $v"
  ncu --set full --clock-control none --import-source on -k regex:"conv_roll_d_kernel" -s 40 -c 2 -o gpurun_out/prof_roll_v$v -f $CMD > gpurun_out/ncu_roll_v$v.log 2>&1
  echo "variant $v ncu rc=$?"
done
ls -la gpurun_out/prof_roll_v*.ncu-rep
