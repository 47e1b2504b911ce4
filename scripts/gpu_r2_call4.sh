#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu ) > gpurun_out/pytest_gpu_r2c.log 2>&1
tail -5 gpurun_out/pytest_gpu_r2c.log
python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --no-extras 2>gpurun_out/bench_r2c.err | tail -1 > gpurun_out/bench_r2c.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2c.json"))
print("value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], d["phases_ms"], d["checks"]["ok"], "launches", d["gpu_launches"])
PY
timeout 900 ncu --set full --import-source on --clock-control none -k regex:trace_face_kernel -c 1 -o gpurun_out/r02_trace_face_v3_c3 -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-extras > gpurun_out/ncu_face.log 2>&1
# launch list of the wide-beam variant (where does the second pass spend its time?)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_wide_beam.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-extras --beam 5e-3 > gpurun_out/ncu_wide.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_default_step.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-extras > gpurun_out/ncu_default.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
