#!/usr/bin/env python
"""Summaries of an ncu report for profiles/ (run here, no GPU needed):

    python scripts/ncu_summarize.py gpurun_out/X.ncu-rep <ray_steps_per_launch> [out_prefix]

writes <out_prefix>_ncu_metrics.csv (selected raw metrics + stall-reason shares), <out_prefix>_opcode_mix.csv (warp
instructions per warp and ray-step by opcode with the average active threads) and prints a cost estimate of the
instruction mix under the operand-bandwidth model measured by scripts/ubench_fp32_pipe.cu."""
import collections, csv, io, re, subprocess, sys

rep, ray_steps = sys.argv[1], float(sys.argv[2])
out = sys.argv[3] if len(sys.argv) > 3 else rep.replace(".ncu-rep", "")
warp_steps = ray_steps / 32.0


def page(name):
    txt = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(txt)))


raw = page("raw")
d = {h: (v, u) for h, u, v in zip(raw[0], raw[1], raw[2])}
KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]
src = page("source")
hdr = src[1]
col = {h: i for i, h in enumerate(hdr)}
rows = [r for r in src[2:] if len(r) == len(hdr)]
ops = collections.defaultdict(lambda: [0.0, 0.0])
stalls = collections.Counter()
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in rows:
    text = r[col["Source"]].strip()
    text = re.sub(r"^@!?U?P\d+\s+", "", text)
    op = text.split()[0].split(".")[0]
    n, t = float(r[col["Instructions Executed"]]), float(r[col["Thread Instructions Executed"]])
    ops[op][0] += n
    ops[op][1] += t
    for h in stall_cols:
        stalls[h] += float(r[col[h]] or 0)
with open(out + "_ncu_metrics.csv", "w") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit", "value"])
    w.writerow(["kernel", "", d["Kernel Name"][0]])
    for k in KEYS:
        if k in d:
            w.writerow([k, d[k][1], d[k][0]])
    w.writerow([])
    w.writerow(["stall_reason", "pct_of_samples"])
    tot = sum(stalls.values())
    for h, v in stalls.most_common(10):
        w.writerow([h, round(100 * v / tot, 1)])
tot_n = sum(v[0] for v in ops.values())
tot_t = sum(v[1] for v in ops.values())
with open(out + "_opcode_mix.csv", "w") as f:
    w = csv.writer(f)
    w.writerow(["opcode", "warp_instructions_per_warp_step", "avg_active_threads"])
    w.writerow(["TOTAL", round(tot_n / warp_steps, 1), round(tot_t / tot_n, 1)])
    for op, (n, t) in sorted(ops.items(), key=lambda kv: -kv[1][0]):
        if n / warp_steps >= 0.05:
            w.writerow([op, round(n / warp_steps, 1), round(t / n, 1)])
per = {op: n / warp_steps for op, (n, t) in ops.items()}
pipe = sum(per.get(o, 0) for o in ("FFMA", "FADD", "FMUL", "IMAD")) + 2 * sum(per.get(o, 0) for o in ("FFMA2", "FADD2", "FMUL2"))
ms = float(d["gpu__time_duration.sum"][0])
print(f"{d['Kernel Name'][0][:60]}: {ms:.2f} ms, {tot_n / warp_steps:.1f} warp-instr per warp-step, {tot_t / tot_n:.1f} active threads")
print(f"FP32-pipe slots per warp-step (1 per scalar, 2 per packed op): {pipe:.1f}")
print("top opcodes:", ", ".join(f"{o} {v:.1f}" for o, v in sorted(per.items(), key=lambda kv: -kv[1])[:16]))
print("stalls:", ", ".join(f"{h[6:]} {100 * v / tot:.0f}%" for h, v in stalls.most_common(8)))
