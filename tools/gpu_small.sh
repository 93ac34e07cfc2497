#!/bin/bash
# small-image workloads with the body-layer kernel forced: PDS_TC_VARIANT 0 (auto) / 16 (1-CTA tiles) / 32 (2-CTA tiles) / 64 (row streaming)
set -u
mkdir -p gpurun_out
for w in cfg1 cfg2; do
for v in "$@"; do
  PDS_TC_VARIANT=$v timeout 600 python bench.py --workload $w --steps 50 --warmup 10 --e2e-iters 50 --no-cpu-baseline --no-hbm-probe > gpurun_out/exp_${w}_v$v.json 2> gpurun_out/exp_${w}_v$v.err
  python - $w $v <<'PY'
import json,sys
w,v=sys.argv[1:3]
d=json.loads(open(f"gpurun_out/exp_{w}_v{v}.json").read().strip().splitlines()[-1])
print(w,'variant',v,'value',round(d['value'],1),'us/step',round(d['ms_per_step']*1e3,1),'e2e',round(d['e2e']['value'],1),'mid avg us',round(d['roofline']['avg_ms']*1e3,2))
PY
done
done
