#!/bin/bash
# Round-2 evidence run on one B200: GPU suite, smoke, reference arm, default bench, the other workloads, ncu launch lists of the
# bench commands (cfg4 shape and the single-image cfg1), ncu --set full captures of the dominant kernels.  Everything lands in
# gpurun_out/; tools/make_profiles.py r02 copies the judged summaries into profiles/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 2000 python -m pytest tests/ -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/BENCH_ref.json 2> gpurun_out/BENCH_ref.err; tail -c 300 gpurun_out/BENCH_ref.json
echo "== bench default"; timeout 900 python bench.py > gpurun_out/BENCH_default.json 2> gpurun_out/BENCH_default.err; echo "rc=$?"; tail -c 300 gpurun_out/BENCH_default.json; tail -3 gpurun_out/BENCH_default.err
for w in cfg1 cfg2 cfg3 cfg2b; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --e2e-iters 100 --no-hbm-probe > gpurun_out/BENCH_$w.json 2> gpurun_out/BENCH_$w.err; echo "$w rc=$?"
done
timeout 600 python bench.py --workload cfg1 --steps 20 --warmup 5 --e2e-iters 100 --no-hbm-probe --no-cpu-baseline --tc-variant 512 > gpurun_out/BENCH_cfg1_per_layer.json 2> gpurun_out/BENCH_cfg1_per_layer.err; echo "cfg1 per-layer rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --workload cfg4 --batch 8 --e2e-iters 1 --no-cpu-baseline --no-hbm-probe"
CMD1="python bench.py --steps 3 --warmup 3 --workload cfg1 --e2e-iters 1 --no-cpu-baseline --no-hbm-probe"
echo "== ncu launch lists"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "rc=$?"; wc -l gpurun_out/launches.csv
$CMD1 > gpurun_out/plain1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_cfg1.csv $CMD1 > gpurun_out/ncu_list1.log 2>&1
echo "rc=$?"; wc -l gpurun_out/launches_cfg1.csv
echo "== ncu full: conv layers (cfg4 shape)"
ncu --set full --clock-control none --import-source on -k regex:"conv_roll_d_kernel|conv_first2_kernel|conv_last_tc_kernel" -s 60 -c 4 -o gpurun_out/prof_conv_layers -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"
echo "== ncu full: stencils (cfg4 shape)"
ncu --set full --clock-control none --import-source on -k regex:"blur_rt_kernel" -s 10 -c 2 -o gpurun_out/prof_blur -f $CMD > gpurun_out/ncu_full_blur.log 2>&1
echo "rc=$?"
echo "== ncu full: im2col first-layer kernel (cross-check, tc_variant bit 15), same shape"
ncu --set full --clock-control none -k regex:"conv_first_tc_kernel" -s 2 -c 1 -o gpurun_out/prof_first_im2col -f $CMD --tc-variant 32768 > gpurun_out/ncu_full_first_im2col.log 2>&1
echo "rc=$?"
echo "== ncu full: chain kernel (cfg1)"
ncu --set full --clock-control none --import-source on -k regex:"conv_chain_kernel" -s 4 -c 1 -o gpurun_out/prof_chain -f $CMD1 > gpurun_out/ncu_full_chain.log 2>&1
echo "rc=$?"
echo "== ncu full: fused pointwise prox kernels (HBM evidence)"
CMD2="python bench.py --steps 1 --warmup 3 --workload cfg2b --e2e-iters 1 --no-cpu-baseline --no-hbm-probe"
ncu --set full --clock-control none --import-source on -k regex:"dual_pw_kernel|primal_pw_kernel|l1ball_kernel" -s 6 -c 3 -o gpurun_out/prof_pointwise -f $CMD2 > gpurun_out/ncu_full2.log 2>&1
echo "rc=$?"; ls -la gpurun_out/*.ncu-rep
python tools/chain_timeline.py 256 256 128 > gpurun_out/timeline_256.txt 2>&1; tail -2 gpurun_out/timeline_256.txt
python tools/gpu_write_bw.py > gpurun_out/hbm_rw.json 2>&1; cat gpurun_out/hbm_rw.json
