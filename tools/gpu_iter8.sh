#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== dncnn"; timeout 600 python -m pytest tests/test_gpu_dncnn.py -m gpu -q --tb=short > gpurun_out/t_dncnn.log 2>&1; echo "rc=$?" >> gpurun_out/t_dncnn.log; tail -4 gpurun_out/t_dncnn.log
for spec in "cfg4 5 16 0" "cfg4 5 16 64" "cfg1 10 64 0" "cfg1 10 64 64"; do
  set -- $spec
  tag="bench_$1_b$3_v$4"
  extra=""; [ "$3" != "0" ] && extra="$extra --batch $3"
  PDS_TC_VARIANT=$4 timeout 900 python bench.py --steps $2 --warmup 3 --workload $1 $extra --no-cpu-baseline > gpurun_out/$tag.json 2> gpurun_out/$tag.err
  python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/{tag}.json').read().strip().splitlines()[-1])
    km={k:round(v['ms']/d['steps'],3) for k,v in d['kernel_ms'].items() if v['launches']}
    print(tag,'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms/step',round(d['ms_per_step'],3),km, d['clocks'])
except Exception as e:
    print(tag,'FAILED',e); print(open(f'gpurun_out/{tag}.err').read()[-1500:])
PY
done
