#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/diag_mallocasync.py > gpurun_out/diag_mallocasync.log 2>&1; cat gpurun_out/diag_mallocasync.log
( time timeout 900 python -m pytest tests -m gpu -q -s --durations=8 ) > gpurun_out/pytest_gpu3.log 2>&1
grep -n "passed\|failed" gpurun_out/pytest_gpu3.log
timeout 600 python scripts/diag_e2e_chunks.py > gpurun_out/diag_e2e_chunks3.log 2>&1
cat gpurun_out/diag_e2e_chunks3.log
( timeout 600 python bench.py --no-cpu ) > gpurun_out/bench_1gpu_r3.json 2> gpurun_out/bench_1gpu_r3.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_1gpu_r3.json").read().strip().splitlines()[-1])
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["phases_ms"], d["e2e"]["trace_kernels_ms_per_step"])
PY
