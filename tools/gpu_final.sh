#!/bin/bash
# Round-end style run on one B200: full GPU suite, smoke, reference arm, default bench, the other workloads,
# ncu launch list of the bench command, ncu --set full captures of the dominant kernels.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/BENCH_ref.json 2> gpurun_out/BENCH_ref.err; tail -c 300 gpurun_out/BENCH_ref.json
echo "== bench default"; timeout 900 python bench.py > gpurun_out/BENCH_default.json 2> gpurun_out/BENCH_default.err; echo "rc=$?"; tail -c 400 gpurun_out/BENCH_default.json; tail -3 gpurun_out/BENCH_default.err
for w in cfg1 cfg2 cfg3 cfg2b cfg5; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --e2e-iters 100 --no-hbm-probe > gpurun_out/BENCH_$w.json 2> gpurun_out/BENCH_$w.err; echo "$w rc=$?"
done
CMD="python bench.py --steps 2 --warmup 3 --workload cfg4 --batch 8 --e2e-iters 1 --no-cpu-baseline --no-hbm-probe"
echo "== ncu launch list"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "rc=$?"; wc -l gpurun_out/launches.csv
echo "== ncu full: conv layers"
ncu --set full --clock-control none --import-source on -k regex:"conv_roll_d_kernel|conv_first_tc_kernel|conv_last_tc_kernel" -s 60 -c 4 -o gpurun_out/prof_conv_layers -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"
echo "== ncu full: fused pointwise prox kernels (HBM evidence)"
CMD2="python bench.py --steps 1 --warmup 3 --workload cfg2b --e2e-iters 1 --no-cpu-baseline --no-hbm-probe"
ncu --set full --clock-control none --import-source on -k regex:"dual_pw_kernel|primal_pw_kernel|l1ball_kernel" -s 6 -c 3 -o gpurun_out/prof_pointwise -f $CMD2 > gpurun_out/ncu_full2.log 2>&1
echo "rc=$?"; ls -la gpurun_out/*.ncu-rep
