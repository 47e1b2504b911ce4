#!/bin/bash
# One GPU-box visit: full GPU test suite, smoke, the default bench line, e2e chunk-size A/B.  Outputs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
( time timeout 1200 python -m pytest tests -m gpu -q -s --durations=12 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1
tail -3 gpurun_out/smoke.log
( time timeout 600 python bench.py ) > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
tail -c 600 gpurun_out/bench_1gpu.json
for chunk in 25000000 50000000; do
  timeout 300 python bench.py --no-cpu --steps 2 --warmup 2 --e2e-chunk $chunk > gpurun_out/bench_chunk_$chunk.json 2> gpurun_out/bench_chunk_$chunk.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_chunk_$chunk.json").read().strip().splitlines()[-1])
print("chunk", $chunk, "value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("phases_ms"))
PY
done
