#!/bin/bash
# Round-2 first pass on one B200: full GPU suite (incl. the reference-recorded BASELINE-shape runs), default bench, single-image configs.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests/ -x -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
grep -E "^BASE_|^UNS_" gpurun_out/pytest_gpu.log | tail -20
echo "== bench default"; timeout 900 python bench.py > gpurun_out/BENCH_default.json 2> gpurun_out/BENCH_default.err; echo "rc=$?"; tail -c 600 gpurun_out/BENCH_default.json; tail -3 gpurun_out/BENCH_default.err
for w in cfg1 cfg2 cfg3; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --e2e-iters 100 --no-hbm-probe > gpurun_out/BENCH_$w.json 2> gpurun_out/BENCH_$w.err; echo "$w rc=$?"; tail -2 gpurun_out/BENCH_$w.err
done
timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 --e2e-iters 10 --no-hbm-probe --no-cpu-baseline > gpurun_out/BENCH_cfg5.json 2> gpurun_out/BENCH_cfg5.err; echo "cfg5 rc=$?"; tail -2 gpurun_out/BENCH_cfg5.err
python - <<'PY'
import json
for w in ["default","cfg1","cfg2","cfg3","cfg5"]:
    try:
        d=json.loads(open(f"gpurun_out/BENCH_{w}.json").read().strip().splitlines()[-1])
        print(w, "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "frac", d["roofline"]["frac"], "parity", d.get("parity"), d["e2e"].get("parts_rank0"))
    except Exception as e:
        print(w, "ERR", e)
PY
