#!/bin/bash
# round 2, call 3: whole GPU suite with the face-coefficient path as the default, default bench (e2e + CPU arms + sub-records), ncu
mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu_r2b.log 2>&1
tail -5 gpurun_out/pytest_gpu_r2b.log
python -m pytest tests/test_gpu_parity_baseline_sizes.py -q -m gpu -s 2>&1 | grep -i "pixel\|oracle\|passed\|failed" > gpurun_out/pytest_sizes_prints.log
( time python bench.py 2>gpurun_out/bench_r2b.err | tail -1 > gpurun_out/bench_r2b.json ) 2>&1 | grep real
tail -3 gpurun_out/bench_r2b.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2b.json"))
print("value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], d["phases_ms"], d["checks"], "launches", d["gpu_launches"])
print("roofline", {k: d["roofline"][k] for k in ("bound", "kernel", "achieved", "peak", "frac", "profile_matches_loaded_kernel")})
print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["phases_ms"])
print("e2e pageable", d["e2e_pageable"] and (d["e2e_pageable"]["value"], d["e2e_pageable"]["ms_per_step"], d["e2e_pageable"]["phases_ms"]))
print("cpu", d["cpu_baseline"], d["cpu_baseline_c"])
for k, v in d["extra"].items():
    print(k, "value %.4g" % v["value"], "ms/step %.2f" % v["ms_per_step"], "kernel %.2f" % v["kernel_ms"], v["phases_ms"], v["checks"]["ok"], v.get("deferred_fraction"))
PY
timeout 900 ncu --set full --import-source on --clock-control none -k regex:trace_face_kernel -c 1 -o gpurun_out/r02_trace_face_v2_c3 -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-extras > gpurun_out/ncu_face.log 2>&1
ls -la gpurun_out/r02_trace_face_v2_c3.ncu-rep
