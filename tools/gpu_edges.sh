#!/bin/bash
# Edge-kernel iteration loop (round 2): operator tests, per-kernel times of the cfg4 step (CUDA events inside bench.py) for the
# tc_variant values given as arguments (0 = default), optionally one ncu --set full capture (NCU=regex).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_dncnn.py -q -x 2>&1 | tail -3
for v in "${@:-0}"; do
  CMD="python bench.py --steps 3 --warmup 3 --workload cfg4 --batch 16 --e2e-iters 1 --no-cpu-baseline --no-hbm-probe --tc-variant $v"
  timeout 300 $CMD > gpurun_out/edges_v$v.json 2> gpurun_out/edges_v$v.err; echo "variant $v rc=$?"
  python - "$v" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/edges_v{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), d["clocks"]["sm_mhz"], {k: round(v["ms"] / max(v["launches"], 1), 4) for k, v in d["kernel_ms"].items()})
PY
done
if [ -n "${NCU:-}" ]; then
  CMD="python bench.py --steps 2 --warmup 3 --workload cfg4 --batch 8 --e2e-iters 1 --no-cpu-baseline --no-hbm-probe --tc-variant ${NCUV:-0}"
  ncu --set full --clock-control none --import-source on -k regex:"$NCU" -s ${NCUS:-8} -c ${NCUC:-4} -o gpurun_out/prof_edges -f $CMD > gpurun_out/ncu_edges.log 2>&1
  echo "ncu rc=$?"; ls -la gpurun_out/prof_edges.ncu-rep
fi
if [ -n "${SMALL:-}" ]; then
  for w in cfg1 cfg3; do
    timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --e2e-iters 100 --no-hbm-probe --no-cpu-baseline > gpurun_out/edges_$w.json 2> gpurun_out/edges_$w.err
    python - "$w" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/edges_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(sys.argv[1], "ms/it", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), {k: round(v["ms"] / max(v["launches"], 1), 4) for k, v in d["kernel_ms"].items()})
PY
  done
fi
