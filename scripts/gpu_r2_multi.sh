#!/bin/bash
# multi-GPU bench as the driver launches it: N = $1
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/gpus_$N.txt
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 2>gpurun_out/bench_${N}gpu.err | tail -1 > gpurun_out/bench_${N}gpu.json ) 2>&1 | grep real
tail -3 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${N}gpu.json"))
print("N", d["n_gpus"], "value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], d["checks"]["ok"], "launches", d["gpu_launches"], "numa", d["config"]["numa_binding"])
print("roofline frac", d["roofline"]["frac"], d["roofline"]["profile_matches_loaded_kernel"])
print("e2e %.4g" % d["e2e"]["value"], "%.2f ms" % d["e2e"]["ms_per_step"], d["e2e"]["phases_ms"][-1] if isinstance(d["e2e"]["phases_ms"], list) else d["e2e"]["phases_ms"])
if d.get("e2e_pageable"): print("e2e pageable %.4g" % d["e2e_pageable"]["value"], "%.2f ms" % d["e2e_pageable"]["ms_per_step"])
for k, v in d["extra"].items():
    print(k, "value %.4g" % v["value"], "ms/step %.2f" % v["ms_per_step"], "kernel %.2f" % v["kernel_ms"], v["checks"]["ok"], "e2e", v.get("e2e", {}).get("value"), v.get("e2e", {}).get("ms_per_step"))
    print("   phases rank0", v["phases_ms"][0] if isinstance(v["phases_ms"], list) else v["phases_ms"])
PY
