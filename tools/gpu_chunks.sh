#!/bin/bash
# cfg4 with different denoiser chunk sizes (images per pass)
set -u
mkdir -p gpurun_out
for c in "$@"; do
  timeout 600 python bench.py --steps 5 --warmup 3 --workload cfg4 --chunk $c --e2e-iters 1 --no-cpu-baseline --no-hbm-probe > gpurun_out/exp_c$c.json 2> gpurun_out/exp_c$c.err
  python - $c <<'PY'
import json,sys
c=sys.argv[1]
d=json.loads(open(f"gpurun_out/exp_c{c}.json").read().strip().splitlines()[-1])
print('chunk',c,'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],2),'mid avg ms',round(d['roofline']['avg_ms'],4),'clocks',d['clocks']['sm_mhz'])
PY
done
