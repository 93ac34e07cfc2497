#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== chain tests"; timeout 600 python -m pytest tests/test_gpu_dncnn.py -x -q -m gpu -k "chain or agree" > gpurun_out/pytest_chain.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_chain.log
for w in cfg1 cfg2 cfg3; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --e2e-iters 100 --no-hbm-probe --no-cpu-baseline > gpurun_out/BENCH_$w.json 2> gpurun_out/BENCH_$w.err; echo "$w rc=$?"; tail -2 gpurun_out/BENCH_$w.err
done
timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 5 --e2e-iters 100 --no-hbm-probe --no-cpu-baseline --tc-variant 128 > gpurun_out/BENCH_cfg2_chain.json 2> gpurun_out/BENCH_cfg2_chain.err
python - <<'PY'
import json
for w in ["cfg1","cfg2","cfg2_chain","cfg3"]:
    try:
        d=json.loads(open(f"gpurun_out/BENCH_{w}.json").read().strip().splitlines()[-1])
        print(w, "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "launches", d["gpu_launches"], {k:(round(v['ms']/max(1,v['launches']),4), v['launches']) for k,v in d['kernel_ms'].items()})
    except Exception as e:
        print(w, "ERR", e)
PY
CMD="python bench.py --workload cfg1 --steps 3 --warmup 3 --e2e-iters 1 --no-cpu-baseline --no-hbm-probe"
ncu --set full --clock-control none --import-source on -k regex:"conv_chain_kernel" -s 4 -c 1 -o gpurun_out/prof_chain2 -f $CMD > gpurun_out/ncu_chain.log 2>&1
echo "ncu rc=$?"
echo "== full"; timeout 1500 python -m pytest tests/ -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.log
