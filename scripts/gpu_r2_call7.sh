#!/bin/bash
mkdir -p gpurun_out
python scripts/sanitizer_target.py > gpurun_out/sanitizer_plain.log 2>&1; tail -1 gpurun_out/sanitizer_plain.log
for tool in memcheck racecheck initcheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitizer_target.py > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "$tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r02_sanitizer_$tool.log | tail -1)"
done
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_memcheck_smoke.log 2>&1
echo "smoke memcheck: $(grep 'ERROR SUMMARY' gpurun_out/r02_sanitizer_memcheck_smoke.log | tail -1)"
( time python bench.py 2>gpurun_out/bench_r2f.err | tail -1 > gpurun_out/bench_r2f.json ) 2>&1 | grep real
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2f.json"))
print("value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], d["checks"]["ok"], "launches", d["gpu_launches"], "roofline", d["roofline"]["frac"], d["roofline"]["profile_matches_loaded_kernel"])
print("e2e %.4g" % d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["phases_ms"], d["e2e"]["pipeline"])
print("e2e pageable %.4g" % d["e2e_pageable"]["value"], d["e2e_pageable"]["ms_per_step"])
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], d["cpu_baseline_c"]["value"])
for k, v in d["extra"].items():
    print(k, "value %.4g" % v["value"], "ms/step %.2f" % v["ms_per_step"], "kernel %.2f" % v["kernel_ms"], v["checks"]["ok"])
PY
