#!/bin/bash
# round 2, call 2: face-coefficient kernel: parity at the benchmarked sizes, bench with / without, ncu of the new kernel
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_parity_baseline_sizes.py -x -q -m gpu -s ) > gpurun_out/pytest_sizes.log 2>&1
tail -5 gpurun_out/pytest_sizes.log
for rep in 1 2; do
python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e 2>gpurun_out/bench_face.err | tail -1 > gpurun_out/bench_face_$rep.json
python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --no-face-grid 2>gpurun_out/bench_noface.err | tail -1 > gpurun_out/bench_noface_$rep.json
done
python - <<'PY'
import json
for n in ("face_1", "noface_1", "face_2", "noface_2"):
    d = json.load(open(f"gpurun_out/bench_{n}.json"))
    print(n, "value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], "kernel_ms %.2f" % d["roofline"]["kernel_ms"], d["phases_ms"], d["histogram_sum"])
PY
timeout 900 ncu --set full --import-source on --clock-control none -k regex:trace_face_kernel -c 1 -o gpurun_out/r02_trace_face_c3 -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_face.log 2>&1
ls -la gpurun_out/r02_trace_face_c3.ncu-rep
