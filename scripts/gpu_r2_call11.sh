#!/bin/bash
# round 2, call 11: the default bench exactly as the driver runs it (N = 1), the reference arm, launch list + detector-kernel ncu
mkdir -p gpurun_out
( time python bench.py 2>gpurun_out/bench_1gpu_r2j.err | tail -1 > gpurun_out/bench_1gpu_r2j.json ) 2>&1 | grep real
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_1gpu_r2j.json"))
print("value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], d["phases_ms"], d["checks"]["ok"], "launches", d["gpu_launches"])
print("roofline", d["roofline"]["frac"], d["roofline"]["profile_matches_loaded_kernel"], "e2e %.4g %.1f ms" % (d["e2e"]["value"], d["e2e"]["ms_per_step"]), d["e2e"]["phases_ms"])
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["cores"])
for k, v in d["extra"].items(): print(k, "%.4g" % v["value"], "%.2f ms" % v["ms_per_step"], v["checks"]["ok"], v["phases_ms"].get("optics+hist") if isinstance(v["phases_ms"], dict) else "")
PY
( time python bench.py --impl reference 2>gpurun_out/bench_ref_r2j.err | tail -1 > gpurun_out/bench_ref_r2j.json ) 2>&1 | grep real
cut -c1-600 gpurun_out/bench_ref_r2j.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_default_step_v2.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-extras > gpurun_out/ncu_launches.log 2>&1
grep -c . gpurun_out/r02_launches_default_step_v2.csv
