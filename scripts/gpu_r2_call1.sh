#!/bin/bash
# round 2, call 1: baseline-size parity tests + A/B of the ld3 / packed-g_w variants
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt
( time python -m pytest tests/test_gpu_parity_baseline_sizes.py -x -q -m gpu -s ) > gpurun_out/pytest_sizes.log 2>&1
tail -5 gpurun_out/pytest_sizes.log
rm -f gpurun_out/ab_variants.txt
bash scripts/ab_variants.sh new nopackw nold3 base
