#!/bin/bash
# Quick pass: denoiser parity + default bench (no CPU leg).  Usage: gpu_quick.sh [extra bench args]
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dncnn.py -m gpu -q --tb=short 2>&1 | tail -4
timeout 900 python bench.py --no-cpu-baseline "$@" > gpurun_out/BENCH_step.json 2> gpurun_out/BENCH_step.err; echo "rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/BENCH_step.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "clocks", d["clocks"])
print("roofline", {k: d["roofline"][k] for k in ("achieved","frac","avg_ms","share_of_step")})
print("kernel_ms", d["kernel_ms"])
PY
tail -3 gpurun_out/BENCH_step.err
