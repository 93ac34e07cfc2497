#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== blur tests"; timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu > gpurun_out/pytest_ops.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_ops.log
for v in 0 8192; do
  timeout 600 python bench.py --steps 5 --warmup 3 --e2e-iters 1 --no-cpu-baseline --no-hbm-probe --tc-variant $v > gpurun_out/BENCH_blur_v$v.json 2> gpurun_out/BENCH_blur_v$v.err
  timeout 600 python bench.py --workload cfg1 --steps 20 --warmup 5 --e2e-iters 10 --no-cpu-baseline --no-hbm-probe --tc-variant $v > gpurun_out/BENCH_blur1_v$v.json 2> gpurun_out/BENCH_blur1_v$v.err
done
python - <<'PY'
import json
for w in ["blur_v0","blur_v8192","blur1_v0","blur1_v8192"]:
    try:
        d=json.loads(open(f"gpurun_out/BENCH_{w}.json").read().strip().splitlines()[-1])
        print(w, "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), {k:(round(v['ms']/max(1,v['launches']),4), v['launches']) for k,v in d['kernel_ms'].items()})
    except Exception as e:
        print(w, "ERR", e)
PY
