#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dncnn.py -m gpu -q -s --tb=short -k "row_streaming" 2>&1 | tail -15
bash tools/gpu_variants.sh 0 128
