#!/bin/bash
# round 2, call 14: staged upload of pageable arrays (tt_h2d_pageable): test, upload rates by thread count, pageable e2e
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -k "pageable or privatised" 2>&1 | tail -3
python - <<'PY'
import os, time, numpy as np, torch
from turbulence_tracing_b200 import _lib
a = np.random.default_rng(0).random(150_000_000)       # 1.2 GB pageable
src = torch.from_numpy(a)
dst = torch.empty_like(src, device="cuda")
torch.cuda.synchronize()
for thr in ("torch", "1", "2", "4", "6", "8"):
    ts = []
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        if thr == "torch":
            dst.copy_(src, non_blocking=True)
        else:
            os.environ["TT_H2D_THREADS"] = thr
            _lib.h2d(dst, src)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print("pageable 1.2 GB, %5s thread(s): %6.1f ms  %5.1f GB/s" % (thr, min(ts) * 1e3, 1.2 / min(ts)))
os.environ.pop("TT_H2D_THREADS", None)
PY
nproc
python bench.py --no-cpu 2>gpurun_out/bench_r2k.err | tail -1 > gpurun_out/bench_r2k.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2k.json"))
print("value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], d["phases_ms"], d["checks"]["ok"])
print("e2e %.4g %.1f ms" % (d["e2e"]["value"], d["e2e"]["ms_per_step"]), d["e2e"]["phases_ms"])
p = d["e2e_pageable"]; print("e2e pageable %.4g %.1f ms" % (p["value"], p["ms_per_step"]), p["phases_ms"], p["pipeline"])
print(d["roofline"]["detector_kernel"])
for k, v in d["extra"].items(): print(k, "%.4g" % v["value"], "%.2f ms" % v["ms_per_step"], v["checks"]["ok"])
PY
