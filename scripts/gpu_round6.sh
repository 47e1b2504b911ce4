#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -s --durations=5 ) > gpurun_out/pytest_gpu6.log 2>&1
grep -n "passed\|failed" gpurun_out/pytest_gpu6.log; grep -n "stretched\|rectilinear" gpurun_out/pytest_gpu6.log | cut -c1-220
timeout 300 python scripts/bench_rectilinear.py > gpurun_out/bench_rectilinear.log 2>&1; cat gpurun_out/bench_rectilinear.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_axes_event -c 1 -o gpurun_out/r01_trace_axes_event_f32 python scripts/bench_rectilinear.py --one > gpurun_out/ncu_rect.log 2>&1; tail -2 gpurun_out/ncu_rect.log
ncu -i gpurun_out/r01_trace_axes_event_f32.ncu-rep --page raw --csv > gpurun_out/r01_trace_axes_event_f32_raw.csv 2>/dev/null; wc -c gpurun_out/r01_trace_axes_event_f32_raw.csv
