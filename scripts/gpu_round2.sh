#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity_c_oracle.py -m gpu -q -s ) > gpurun_out/pytest_gpu2.log 2>&1
tail -3 gpurun_out/pytest_gpu2.log
timeout 600 python scripts/diag_e2e_chunks.py > gpurun_out/diag_e2e_chunks.log 2>&1
cat gpurun_out/diag_e2e_chunks.log
