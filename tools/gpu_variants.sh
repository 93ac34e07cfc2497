#!/bin/bash
# Perf experiments on the body-layer kernel: --tc-variant values (pds_debug_set_tc_variant) given as arguments, cfg4 shape, 16 images.
set -u
mkdir -p gpurun_out
for v in "$@"; do
  timeout 600 python bench.py --steps 3 --warmup 3 --workload cfg4 --batch 16 --e2e-iters 1 --no-cpu-baseline --no-hbm-probe --tc-variant $v > gpurun_out/exp_v$v.json 2> gpurun_out/exp_v$v.err
  python - $v <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/exp_v{v}.json").read().strip().splitlines()[-1])
    print('variant',v,'ms/step',round(d['ms_per_step'],2),'mid avg ms',round(d['roofline']['avg_ms'],4),'clocks',d['clocks']['sm_mhz'], 'psnr', round(d['quality']['final_psnr_mean'],2))
except Exception as e:
    print('variant',v,'failed',e); print(open(f"gpurun_out/exp_v{v}.err").read()[-500:])
PY
done
