#!/bin/bash
set -u
mkdir -p gpurun_out
for spec in "cfg2b 5 0 0"; do
  set -- $spec
  tag="bench_$1_b$3_c$4"
  timeout 900 python bench.py --steps $2 --warmup 3 --workload $1 --no-cpu-baseline > gpurun_out/$tag.json 2> gpurun_out/$tag.err
  python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/{tag}.json').read().strip().splitlines()[-1])
    km={k:round(v['ms']/d['steps'],3) for k,v in d['kernel_ms'].items() if v['launches']}
    print(tag,'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),km, d['roofline_hbm'])
except Exception as e:
    print(tag,'FAILED',e); print(open(f'gpurun_out/{tag}.err').read()[-1500:])
PY
done
CMD="python bench.py --steps 1 --warmup 3 --workload cfg4 --batch 8 --e2e-iters 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"conv_first_kernel|blur_rt_kernel|conv_tc_kernel<16>" -s 6 -c 4 -o gpurun_out/prof_thin $CMD > gpurun_out/ncu_thin.log 2>&1
echo "rc=$?"; ls -la gpurun_out/prof_thin.ncu-rep
