#!/bin/bash
# BASELINE configs[4] on the machine it is defined for: the grid search (256 gray 256x256 images x 8 grid points = 2048 items,
# fixed total) sharded over N = 1, 2, 4, 8 GPUs of one box through main.grid_search, PSNR rows gathered over NCCL.
set -u
mkdir -p gpurun_out
ITERS=${ITERS:-20}
for n in 8 4 2 1; do
  if [ $n -eq 1 ]; then
    timeout 900 python bench.py --workload cfg5 --gpus 1 --steps 5 --warmup 3 --e2e-iters $ITERS --no-hbm-probe --no-cpu-baseline > gpurun_out/BENCH_cfg5_${n}gpu.json 2> gpurun_out/BENCH_cfg5_${n}gpu.err
  else
    NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --workload cfg5 --gpus $n --steps 5 --warmup 3 --e2e-iters $ITERS --no-hbm-probe --no-cpu-baseline > gpurun_out/BENCH_cfg5_${n}gpu.json 2> gpurun_out/BENCH_cfg5_${n}gpu.err
  fi
  echo "N=$n rc=$?"; tail -2 gpurun_out/BENCH_cfg5_${n}gpu.err
done
python - <<'PY'
import json
base=None
for n in (1,2,4,8):
    try:
        d=json.loads(open(f"gpurun_out/BENCH_cfg5_{n}gpu.json").read().strip().splitlines()[-1])
        if n==1: base=d
        print(n, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "s/call", round(d["e2e"]["seconds_per_call"],3), d["e2e"]["parts_rank0"],
              "eff(value)", round(d["value"]/(n*base["value"]),3) if base else None, "eff(e2e)", round(d["e2e"]["value"]/(n*base["e2e"]["value"]),3) if base else None)
    except Exception as e:
        print(n, "ERR", e)
PY
