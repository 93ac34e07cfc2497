#!/bin/bash
# bench lines + ncu launch list + one full capture of the top kernel.  Logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
W=${1:-cfg4}
echo "== probe"; timeout 300 python -m pytest tests/test_gpu_tcgen05_probe.py -m gpu -q -s --tb=short > gpurun_out/t_probe.log 2>&1; tail -3 gpurun_out/t_probe.log
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.json
echo "== bench $W"; timeout 900 python bench.py --steps 5 --warmup 3 --workload $W > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; echo "rc=$?"; tail -c 3000 gpurun_out/bench_$W.json; tail -5 gpurun_out/bench_$W.err
echo "== bench cfg1"; timeout 600 python bench.py --steps 20 --warmup 5 --workload cfg1 --no-cpu-baseline > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err; tail -c 2500 gpurun_out/bench_cfg1.json
echo "== bench cfg1 batch 64"; timeout 600 python bench.py --steps 10 --warmup 3 --workload cfg1 --batch 64 --no-cpu-baseline > gpurun_out/bench_cfg1_b64.json 2> gpurun_out/bench_cfg1_b64.err; tail -c 2500 gpurun_out/bench_cfg1_b64.json
echo "== bench cfg2"; timeout 600 python bench.py --steps 10 --warmup 3 --workload cfg2 --no-cpu-baseline > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -c 2500 gpurun_out/bench_cfg2.json
echo "== bench simt"; timeout 900 python bench.py --steps 2 --warmup 3 --workload $W --batch 8 --engine simt --no-cpu-baseline > gpurun_out/bench_simt.json 2> gpurun_out/bench_simt.err; tail -c 1500 gpurun_out/bench_simt.json
