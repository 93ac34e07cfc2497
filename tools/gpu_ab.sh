#!/bin/bash
# Steady-state A/B of --tc-variant values (pds_debug_set_tc_variant) on the default workload (cfg4, 64 images): gpu_ab.sh V1 V2 ...
set -u
mkdir -p gpurun_out
for v in "$@"; do
  timeout 600 python bench.py --steps 6 --warmup 3 --e2e-iters 1 --no-cpu-baseline --no-hbm-probe --tc-variant $v > gpurun_out/ab_v$v.json 2> gpurun_out/ab_v$v.err
  python - $v <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/ab_v{v}.json").read().strip().splitlines()[-1])
    print('variant',v,'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],2),'mid avg ms',round(d['roofline']['avg_ms'],4),'clocks',d['clocks'], 'psnr', round(d['quality']['final_psnr_mean'],3))
except Exception as e:
    print('variant',v,'failed',e); print(open(f"gpurun_out/ab_v{v}.err").read()[-500:])
PY
done
