#!/usr/bin/env python
"""Refresh one entry of profiles/traffic.json from an ncu report of the trace kernel (run here, no GPU needed):

    python scripts/update_traffic.py gpurun_out/X.ncu-rep <key> <ray_steps_per_launch> <mangled-kernel-substring> [profiles prefix]

key = "<workload>/<dtype>/<faces|variant>" as bench.py looks it up.  The entry is stamped with the sha256 of the kernel's
SASS in the CURRENT turbulence_tracing_b200/libtt_b200.so (bench.kernel_sass_sha), which bench.py verifies against the
library it loaded: refresh the entry in the same commit as any change to the kernel."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

rep, key, ray_steps, mangled = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
prefix = sys.argv[5] if len(sys.argv) > 5 else None
if prefix:
    subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summarize.py"), rep, str(ray_steps), prefix], check=True)
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
raw = list(csv.reader(io.StringIO(txt)))
d = {h: v for h, v in zip(raw[0], raw[2])}
f = lambda k: float(d[k].replace(",", ""))
mix = {}
if prefix:
    for r in csv.DictReader(open(prefix + "_opcode_mix.csv")):
        mix[r["opcode"]] = float(r["warp_instructions_per_warp_step"])
slots = (sum(mix.get(o, 0) for o in ("FFMA", "FADD", "FMUL", "IMAD")) + 2 * sum(mix.get(o, 0) for o in ("FFMA2", "FADD2", "FMUL2"))) if mix else None
unit = {h: u for h, u in zip(raw[0], raw[1])}
to_bytes = lambda k: f(k) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[unit[k]]
e = {
    "kernel": d["Kernel Name"], "sass_sha256": bench.kernel_sass_sha(mangled), "kernel_mangled_substring": mangled,
    "dram_bytes_per_launch": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
    "read": to_bytes("dram__bytes_read.sum"), "write": to_bytes("dram__bytes_write.sum"),
    "ncu_kernel_ms": f("gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}[unit["gpu__time_duration.sum"]],
    "fma_pipe_busy_pct": f("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    "issue_slots_busy_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "l1_hit_pct": f("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": f("lts__t_sector_hit_rate.pct"),
    "dram_pct_of_peak": f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    "registers": int(f("launch__registers_per_thread")),
    "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "eligible_warps_per_scheduler": f("smsp__warps_eligible.avg.per_cycle_active"),
    "warp_instructions_per_warp_step": mix.get("TOTAL"),
    "fma_pipe_slots_per_warp_step": slots,
    "fma_pipe_slots_note": "FP32 (FMA) pipe issue slots per warp and ray-step from the opcode mix of this capture: FFMA + FADD + FMUL + IMAD "
                           "+ 2 x (FFMA2 + FADD2 + FMUL2) -- a packed FP32x2 instruction holds the pipe for two cycles",
    "ray_steps_per_launch": ray_steps,
    "source": (os.path.relpath(prefix, ROOT) + "_ncu_metrics.csv / _opcode_mix.csv" if prefix else os.path.basename(rep)) + " (ncu --set full --clock-control none)",
}
path = os.path.join(ROOT, "profiles", "traffic.json")
j = json.load(open(path))
j["entries"][key] = e
j["note"] = ("per-launch ncu figures of the trace kernels, one `ncu --set full` capture per entry (workload/dtype/path), B200; entries written by "
             "scripts/update_traffic.py carry the sha256 of the kernel's SASS, which bench.py checks against the loaded library")
json.dump(j, open(path, "w"), indent=1)
print(json.dumps(e, indent=1))
