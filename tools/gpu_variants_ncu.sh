#!/bin/bash
# ncu counters (operand wavefronts, tensor-pipe activity, cycles) of the body-layer kernel for --tc-variant values (pds_debug_set_tc_variant).
set -u
mkdir -p gpurun_out
M=l1tex__data_pipe_tc_wavefronts_mem_shared.sum,sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg,sm__cycles_elapsed.avg,gpu__time_duration.sum,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum
for v in "$@"; do
  ncu --metrics $M --clock-control none -k regex:"conv_tc2_kernel|conv_roll_kernel" -s 20 -c 1 --csv --log-file gpurun_out/ncu_v$v.csv \
    python bench.py --steps 2 --warmup 3 --workload cfg4 --batch 8 --e2e-iters 1 --no-cpu-baseline --no-hbm-probe --tc-variant $v > gpurun_out/ncu_v$v.log 2>&1
  echo "variant $v rc=$?"; grep -v "^==" gpurun_out/ncu_v$v.csv | cut -d, -f5,13- | tail -9
done
