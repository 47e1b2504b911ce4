#!/bin/bash
# final check of a round on one B200: GPU tests, smoke, bench (+ optional ncu launch list)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu_final.log 2>&1; grep -n "passed\|failed" gpurun_out/pytest_gpu_final.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke_final.log 2>&1; tail -2 gpurun_out/smoke_final.log
( timeout 600 python bench.py ) > gpurun_out/bench_1gpu_final.json 2> gpurun_out/bench_1gpu_final.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_1gpu_final.json").read().strip().splitlines()[-1])
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["phases_ms"], d["e2e"]["trace_launches_per_step"], d["e2e"]["pipeline"])
PY
if [ "$1" = "ncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_default_full_ncu.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
  wc -l gpurun_out/launches_default_full_ncu.csv
fi
