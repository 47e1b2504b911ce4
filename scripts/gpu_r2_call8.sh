#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu ) > gpurun_out/pytest_gpu_r2g.log 2>&1
tail -4 gpurun_out/pytest_gpu_r2g.log
python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --no-extras 2>gpurun_out/bench_r2g.err | tail -1 > gpurun_out/bench_r2g.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2g.json"))
print("value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], d["phases_ms"], d["checks"]["ok"], "launches", d["gpu_launches"])
PY
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:optics_hist -c 4 --csv --log-file gpurun_out/r02_optics_hist_ncu.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-extras > gpurun_out/ncu_hist.log 2>&1
grep -i "optics_hist" gpurun_out/r02_optics_hist_ncu.csv | cut -d, -f5,12-15 | tail -6
