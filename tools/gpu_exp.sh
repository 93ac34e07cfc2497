#!/bin/bash
set -u
mkdir -p gpurun_out
for v in 0 2 6 8 14 16 22; do
  PDS_TC_VARIANT=$v timeout 600 python bench.py --steps 3 --warmup 3 --workload cfg4 --batch 16 --no-cpu-baseline > gpurun_out/exp_v$v.json 2> gpurun_out/exp_v$v.err
  python - $v <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/exp_v{v}.json').read().strip().splitlines()[-1])
    km={k:round(v_['ms']/d['steps'],3) for k,v_ in d['kernel_ms'].items() if v_['launches']}
    print('variant',v,'ms/step',round(d['ms_per_step'],2),'mid avg ms',round(d['roofline']['avg_ms'],3),'TF',round(d['roofline']['achieved'],1),km,d['clocks'])
except Exception as e:
    print(v,'FAILED',e); print(open(f'gpurun_out/exp_v{v}.err').read()[-800:])
PY
done
