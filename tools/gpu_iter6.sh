#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== dncnn (2-CTA engine)"; timeout 300 python -m pytest tests/test_gpu_dncnn.py -m gpu -q -x --tb=short > gpurun_out/t_dncnn.log 2>&1; echo "rc=$?" >> gpurun_out/t_dncnn.log; tail -15 gpurun_out/t_dncnn.log
if grep -q "rc=0" gpurun_out/t_dncnn.log; then
echo "== loops"; timeout 900 python -m pytest tests/test_gpu_loops.py -m gpu -q --tb=short > gpurun_out/t_loops.log 2>&1; echo "rc=$?" >> gpurun_out/t_loops.log; tail -7 gpurun_out/t_loops.log
for spec in "cfg4 5 0 0 0" "cfg4 5 0 0 16" "cfg1 20 0 0 0" "cfg1 10 64 0 0" "cfg2 10 0 0 0"; do
  set -- $spec
  tag="bench_$1_b$3_v$5"
  extra=""; [ "$3" != "0" ] && extra="$extra --batch $3"
  PDS_TC_VARIANT=$5 timeout 900 python bench.py --steps $2 --warmup 3 --workload $1 $extra --no-cpu-baseline > gpurun_out/$tag.json 2> gpurun_out/$tag.err
  python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/{tag}.json').read().strip().splitlines()[-1])
    km={k:round(v['ms']/d['steps'],3) for k,v in d['kernel_ms'].items() if v['launches']}
    print(tag,'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms/step',round(d['ms_per_step'],3),'mid TF',round(d['roofline']['achieved'],1),'frac',round(d['roofline']['frac'],3),km, d['clocks'])
except Exception as e:
    print(tag,'FAILED',e); print(open(f'gpurun_out/{tag}.err').read()[-1500:])
PY
done
fi
