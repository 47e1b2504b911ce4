#!/usr/bin/env python
"""Static view of the production trace kernel's SASS: size of the marching loop and its opcode mix.

    python scripts/sass_loop.py [build/obj/trace_event.o] [mangled-name-substring]
The loop is taken as the span between the last backward branch and its target."""
import collections, re, subprocess, sys
obj = sys.argv[1] if len(sys.argv) > 1 else "build/obj/trace_event.o"
key = sys.argv[2] if len(sys.argv) > 2 else "trace_event_kernel_f32x2ILb1ELb0ELb1ELb0E"
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
fn, cur = {}, None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); fn[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur:
        fn[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in fn.items():
    if key not in name:
        continue
    back = [(a, int(re.search(r"0x([0-9a-f]+)", t).group(1), 16)) for a, t in ins if re.search(r"\bBRA\b", t) and re.search(r"0x([0-9a-f]+)", t) and int(re.search(r"0x([0-9a-f]+)", t).group(1), 16) < a]
    # the marching loop = the backward branch with the longest span
    a1, a0 = max(back, key=lambda b: b[0] - b[1])
    loop = [t for a, t in ins if a0 <= a <= a1]
    ops = collections.Counter(re.sub(r"^@!?U?P\d\s+", "", t).split()[0].split(".")[0] for t in loop)
    print(name[:70], "total", len(ins), "loop", len(loop), f"[{a0:#x}-{a1:#x}]")
    print("  ", ", ".join(f"{k} {v}" for k, v in ops.most_common(24)))
