#!/bin/bash
# One GPU-box pass: hardware probes, parity tests, a short bench.  Logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import torch; print(torch.cuda.get_device_name(0))" >> gpurun_out/gpu.txt 2>&1
echo "== probes";  timeout 300 python -m pytest tests/test_gpu_tcgen05_probe.py -m gpu -q -s --tb=short > gpurun_out/t_probe.log 2>&1; echo "rc=$?" >> gpurun_out/t_probe.log; tail -5 gpurun_out/t_probe.log
echo "== ops";     timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short > gpurun_out/t_ops.log 2>&1; echo "rc=$?" >> gpurun_out/t_ops.log; tail -15 gpurun_out/t_ops.log
echo "== dncnn";   timeout 600 python -m pytest tests/test_gpu_dncnn.py -m gpu -q -s --tb=short > gpurun_out/t_dncnn.log 2>&1; echo "rc=$?" >> gpurun_out/t_dncnn.log; tail -25 gpurun_out/t_dncnn.log
echo "== loops";   timeout 1500 python -m pytest tests/test_gpu_loops.py -m gpu -q -s --tb=short > gpurun_out/t_loops.log 2>&1; echo "rc=$?" >> gpurun_out/t_loops.log; tail -25 gpurun_out/t_loops.log
echo "== smoke";   timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
