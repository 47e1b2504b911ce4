#!/bin/bash
# round 2, final records at 1 GPU: GPU suite, smoke(), the default bench as the driver runs it, the reference arm, launch list
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu ) > gpurun_out/pytest_gpu_r2_final.log 2>&1; tail -4 gpurun_out/pytest_gpu_r2_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time python bench.py 2>gpurun_out/bench_1gpu_r2_final.err | tail -1 > gpurun_out/bench_1gpu_r2_final.json ) 2>&1 | grep real
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_1gpu_r2_final.json"))
print("value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], d["phases_ms"], d["checks"]["ok"], "launches", d["gpu_launches"])
print("roofline", d["roofline"]["frac"], d["roofline"]["profile_matches_loaded_kernel"], d["roofline"]["detector_kernel"]["frac"])
print("e2e %.4g %.1f ms" % (d["e2e"]["value"], d["e2e"]["ms_per_step"]), d["e2e"]["phases_ms"])
p = d["e2e_pageable"]; print("e2e pageable %.4g %.1f ms" % (p["value"], p["ms_per_step"]), p["phases_ms"], p["pipeline"])
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["cores"], "c", d["cpu_baseline_c"]["value"])
for k, v in d["extra"].items(): print(k, "%.4g" % v["value"], "%.2f ms" % v["ms_per_step"], v["checks"]["ok"])
PY
( time python bench.py --impl reference 2>gpurun_out/bench_ref_r2_final.err | tail -1 > gpurun_out/bench_ref_r2_final.json ) 2>&1 | grep real
cut -c1-300 gpurun_out/bench_ref_r2_final.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_default_step_v3.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-extras > gpurun_out/ncu_launches.log 2>&1
grep -c . gpurun_out/r02_launches_default_step_v3.csv
