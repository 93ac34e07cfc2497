#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== dncnn"; timeout 600 python -m pytest tests/test_gpu_dncnn.py -m gpu -q --tb=short > gpurun_out/t_dncnn.log 2>&1; echo "rc=$?" >> gpurun_out/t_dncnn.log; tail -4 gpurun_out/t_dncnn.log
echo "== loops(long)"; timeout 900 python -m pytest tests/test_gpu_loops.py -m gpu -q -s --tb=short -k "full_iteration" > gpurun_out/t_loops.log 2>&1; echo "rc=$?" >> gpurun_out/t_loops.log; tail -7 gpurun_out/t_loops.log
bash tools/gpu_exp.sh
