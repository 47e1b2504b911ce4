#!/bin/bash
mkdir -p gpurun_out
./build/ubench_fp32_pipe 1965 > gpurun_out/r02_ubench_fp32_pipe.txt 2>&1; cat gpurun_out/r02_ubench_fp32_pipe.txt
( time timeout 1200 python -m pytest tests/test_gpu_parity_baseline_sizes.py -m gpu -x -q -s 2>&1 | tail -40 ) > gpurun_out/pytest_sizes.log 2>&1; tail -25 gpurun_out/pytest_sizes.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_event_kernel_f32x2 -s 1 -c 1 -o gpurun_out/r02_trace_v5_c3 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_v5.log 2>&1; tail -3 gpurun_out/ncu_v5.log; ls -la gpurun_out/*.ncu-rep
