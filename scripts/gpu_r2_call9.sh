#!/bin/bash
# round 2, call 9: GPU suite with the 16-bit privatised detector kernel, quick bench, ncu of the detector kernel, launch-shape A/B of the face kernel
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu ) > gpurun_out/pytest_gpu_r2h.log 2>&1
tail -4 gpurun_out/pytest_gpu_r2h.log
python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --no-extras 2>gpurun_out/bench_r2h.err | tail -1 > gpurun_out/bench_r2h.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2h.json"))
print("value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], d["phases_ms"], d["checks"]["ok"], "launches", d["gpu_launches"])
PY
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"optics_hist|SortOnesweep|face_grid|calc_dndr|morton" -c 24 --csv --log-file gpurun_out/r02_small_kernels_ncu.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-extras > gpurun_out/ncu_small.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r02_small_kernels_ncu.csv")))
hdr = None; agg = collections.OrderedDict()
for r in rows:
    if r and r[0] == "ID": hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        agg.setdefault(d["Kernel Name"][:48], {}).setdefault(d["Metric Name"], []).append(float(d["Metric Value"]))
for k, m in agg.items():
    print(k, {a: round(sum(b) / len(b) / 1e6, 3) for a, b in m.items()})
PY
rm -f gpurun_out/ab_variants.txt
TT_BENCH_EXTRA="--no-extras" bash scripts/ab_variants.sh f_new f_b64 f_b96 f_mb4 f_mb6 f_b256 f_b192 f_nofast
