#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu ) > gpurun_out/pytest_gpu_r2e.log 2>&1
tail -5 gpurun_out/pytest_gpu_r2e.log
for extra in "" "--beam 5e-3" "--workload c5 --steps 2 --warmup 1" "--workload c2"; do
python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --no-extras $extra 2>gpurun_out/bench_r2e.err | tail -1 > gpurun_out/bench_r2e.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2e.json"))
print(d["config"]["cube"], "value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], d["phases_ms"], d["checks"]["ok"], "launches", d["gpu_launches"])
PY
done
timeout 900 ncu --set full --import-source on --clock-control none -k regex:trace_face_kernel -c 1 -o gpurun_out/r02_trace_face_v6_c3 -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-extras > gpurun_out/ncu_face.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -1
