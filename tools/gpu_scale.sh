#!/bin/bash
# scaling run: bench at N GPUs of one box (torchrun), same command line the driver uses
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/SCALE_$N.json 2> gpurun_out/SCALE_$N.err
echo "rc=$?"; tail -c 1200 gpurun_out/SCALE_$N.json; tail -5 gpurun_out/SCALE_$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/SCALE_ref_$N.json 2> gpurun_out/SCALE_ref_$N.err
echo "ref rc=$?"; tail -c 300 gpurun_out/SCALE_ref_$N.json
