#!/bin/bash
# round 2, call 10: rebase loop of the face kernel -- GPU suite, A/B against the v4 loop, ncu full capture
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu -x ) > gpurun_out/pytest_gpu_r2i.log 2>&1
tail -4 gpurun_out/pytest_gpu_r2i.log
rm -f gpurun_out/ab_variants.txt
TT_BENCH_EXTRA="--no-extras" bash scripts/ab_variants.sh f_new f_norebase f_b64 f_mb6
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_face_kernel -s 1 -c 1 -o gpurun_out/r02_trace_face_v7_c3 -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-extras > gpurun_out/ncu_v7.log 2>&1
ls -la gpurun_out/r02_trace_face_v7_c3.ncu-rep
