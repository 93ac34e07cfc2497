#!/bin/bash
# One development pass on the GPU box: denoiser + loop parity, then the default bench without the CPU leg.
set -u
mkdir -p gpurun_out
echo "== dncnn";   timeout 600 python -m pytest tests/test_gpu_dncnn.py -m gpu -q -s --tb=short > gpurun_out/t_dncnn.log 2>&1; echo "rc=$?" >> gpurun_out/t_dncnn.log; grep -v "^$" gpurun_out/t_dncnn.log | tail -40
echo "== loops";   timeout 1500 python -m pytest tests/test_gpu_loops.py -m gpu -q -s --tb=short > gpurun_out/t_loops.log 2>&1; echo "rc=$?" >> gpurun_out/t_loops.log; grep "rel_l2\|passed\|failed\|rc=\|Error\|error" gpurun_out/t_loops.log | tail -70
echo "== bench"; timeout 900 python bench.py --no-cpu-baseline > gpurun_out/BENCH_step.json 2> gpurun_out/BENCH_step.err; echo "rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/BENCH_step.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "clocks", d["clocks"])
print("roofline", {k: d["roofline"][k] for k in ("achieved","frac","avg_ms","share_of_step")})
print("kernel_ms", d["kernel_ms"])
PY
tail -3 gpurun_out/BENCH_step.err
