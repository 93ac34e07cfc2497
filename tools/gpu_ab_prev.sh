#!/bin/bash
# A/B of the current library against pnp-pds_b200/libpnp_pds_prev.so (a build of an earlier commit), interleaved, same board.
set -u
mkdir -p gpurun_out
cd pnp-pds_b200 && cp libpnp_pds.so libpnp_pds_cur.so && cd ..
for rep in 1 2 3; do
  for which in prev cur; do
    cp pnp-pds_b200/libpnp_pds_$which.so pnp-pds_b200/libpnp_pds.so
    timeout 600 python bench.py --steps 6 --warmup 3 --e2e-iters 1 --no-cpu-baseline --no-hbm-probe "$@" > gpurun_out/abp_$which.json 2> gpurun_out/abp_$which.err
    python - $which <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/abp_{v}.json").read().strip().splitlines()[-1])
    print(v,'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],2),'mid avg ms',round(d['roofline']['avg_ms'],4),'clocks',d['clocks']['sm_mhz'],'psnr',round(d['quality']['final_psnr_mean'],4))
except Exception as e:
    print(v,'failed',e); print(open(f"gpurun_out/abp_{v}.err").read()[-600:])
PY
  done
done
cp pnp-pds_b200/libpnp_pds_cur.so pnp-pds_b200/libpnp_pds.so
