#!/bin/bash
# Operator / loop / driver parity tests, then the single-image workloads (cfg1, cfg2, cfg3) without the CPU leg.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_loops.py tests/test_gpu_driver.py -m gpu -q -x --tb=short 2>&1 | tail -4
for w in cfg1 cfg2 cfg3; do
  timeout 200 python bench.py --workload $w --steps 20 --warmup 5 --e2e-iters 100 --no-hbm-probe --no-cpu-baseline > gpurun_out/S_$w.json 2> gpurun_out/S_$w.err
  python - $w <<'PY'
import json, sys
w = sys.argv[1]
d = json.loads(open(f"gpurun_out/S_{w}.json").read().strip().splitlines()[-1])
print(w, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 4),
      "primal_ms", round(d["stencil_kernels"]["primal_ms"], 4), "dual_ms", round(d["stencil_kernels"]["dual_ms"], 4))
PY
done
