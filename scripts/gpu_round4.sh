#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -s --durations=8 ) > gpurun_out/pytest_gpu4.log 2>&1
grep -n "passed\|failed" gpurun_out/pytest_gpu4.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke4.log 2>&1; tail -4 gpurun_out/smoke4.log
( timeout 600 python bench.py ) > gpurun_out/bench_1gpu_r4.json 2> gpurun_out/bench_1gpu_r4.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_1gpu_r4.json").read().strip().splitlines()[-1])
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["phases_ms"], d["e2e"]["trace_kernels_ms_per_step"], d["e2e"]["trace_launches_per_step"])
print(d["cpu_baseline"]["value"], d["cpu_baseline_c"]["value"], d["clocks"])
PY
( timeout 300 python bench.py --impl reference --steps 1 --warmup 1 ) > gpurun_out/bench_ref_r4.json 2> gpurun_out/bench_ref_r4.err; tail -c 900 gpurun_out/bench_ref_r4.json
