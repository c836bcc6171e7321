#!/usr/bin/env python3
"""Warp-stall samples of one kernel of an .ncu-rep (ncu --set full --import-source on), joined with nvdisasm's line table of the SAME build
(-lineinfo) and aggregated by source region of csrc/env_step_core.cuh (a region starts at a function or at a `// ----` phase comment).
Usage: tools/ncu_stalls_by_line.py report.ncu-rep build/env_step.cu.o step_kernel_half [--lines 30]"""
import csv, io, re, subprocess, sys, tempfile, os, collections
rep, obj, kname = sys.argv[1:4]
nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 25
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# instructions of the kernel with the innermost source line in force
ins, on, cur = [], False, ("?", 0)
for ln in dis:
    if ln.startswith("\t.section\t.text."):
        on = kname in ln
        continue
    if not on:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?) ;", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2), cur))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
body = rows[2:]
base = int(body[0][0], 16)
by_off = {off: (off, txt, cur) for off, txt, cur in ins}
last = ins[0]
joined = []
for r in body:       # join on the offset inside the kernel; instructions nvdisasm prints in another form inherit the previous line
    last = by_off.get(int(r[0], 16) - base, last)
    joined.append(last)
ins = joined
ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "(Not Issued)" not in h]
src = open(os.environ.get("GO2_STALL_SRC") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "go2_rl_gym_b200", "csrc", "env_step_core.cuh")).read().splitlines()   # GO2_STALL_SRC: the source of the build the report was taken from
def region(line):
    for k in range(line - 1, -1, -1):
        s = src[k]
        if re.match(r"\s*// ---- ", s) or re.match(r"GO2_HD \w+ \w+\(", s):
            return f"{k + 1}: {s.strip()[:90]}"
    return "?"
by_line, by_reg = collections.Counter(), collections.Counter()
st_line, st_reg = collections.defaultdict(collections.Counter), collections.defaultdict(collections.Counter)
ex_reg = collections.Counter()
tot = 0
for r, (off, txt, (f, l)) in zip(body, ins):
    n = int(r[ci["# Samples"]] or 0)
    tot += n
    key = (f, l)
    reg = region(l) if f == "env_step_core.cuh" else f
    by_line[key] += n; by_reg[reg] += n
    ex_reg[reg] += int(r[ci["Instructions Executed"]] or 0)
    for s in stalls:
        v = int(r[ci[s]] or 0)
        if v:
            st_line[key][s] += v; st_reg[reg][s] += v
print(f"kernel {kname}: {len(ins)} SASS instructions, {tot} warp-stall samples")
print("\n== by source region (samples, share, warp instructions executed, top stall reasons)")
for reg, n in by_reg.most_common(40):
    top = ", ".join(f"{s[6:]} {100 * v / max(n, 1):.0f}%" for s, v in st_reg[reg].most_common(4))
    print(f"{n:8d} {100 * n / tot:5.1f}%  {ex_reg[reg]:10d}  {reg:100s} {top}")
print("\n== by source line")
for (f, l), n in by_line.most_common(nlines):
    top = ", ".join(f"{s[6:]} {100 * v / max(n, 1):.0f}%" for s, v in st_line[(f, l)].most_common(3))
    text = src[l - 1].strip()[:100] if f == "env_step_core.cuh" and l <= len(src) else ""
    print(f"{n:8d} {100 * n / tot:5.1f}%  {f}:{l:<5d} {text:100s} {top}")
