#!/usr/bin/env python
"""Per-source-line profile of one kernel: joins an ncu report's SASS page (instructions executed, stall
samples per instruction address) with nvdisasm's line table of the cubin inside libpsim.so.

    line_profile.py report.ncu-rep <mangled-kernel-substring> <units-per-launch> [min_share]

Prints, per source line, warp-instructions per unit (e.g. per tile) and the share of stall samples."""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, kern, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
min_share = float(sys.argv[4]) if len(sys.argv) > 4 else 0.004
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "parallel-particle-simulation_b200", "csrc", "build", "libpsim.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if (sys.argv[5] if len(sys.argv) > 5 else "tiled") in f][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
line_of, cur, infn = {}, None, False
for l in sass.splitlines():
    if l.startswith(".text."):
        infn = kern in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m:
        line_of[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]
data = [r for r in rows[2:] if len(r) >= len(h)]
ai, si, ii = h.index("Address"), h.index("# Samples"), h.index("Instructions Executed")
base = int(data[0][ai], 16)
agg = collections.OrderedDict()
S = sum(int(r[si]) for r in data)
for r in data:
    key = line_of.get(int(r[ai], 16) - base, ("?", 0))
    a = agg.setdefault(key, [0, 0, 0])
    a[0] += int(r[ii]); a[1] += int(r[si]); a[2] += 1
src = {}
print(f"total warp-instr/unit {sum(a[0] for a in agg.values()) / units:.1f}, samples {S}")
print(" file:line            warp-instr/unit  samples%  sass  source")
for (f, ln), (ins, smp, n) in sorted(agg.items(), key=lambda kv: (kv[0][0], kv[0][1])):
    if ins / units < 3 and smp < min_share * S:
        continue
    if f not in src:
        p = os.path.join(root, "parallel-particle-simulation_b200", "csrc", f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = src[f][ln - 1].strip()[:90] if 0 < ln <= len(src[f]) else ""
    print(f" {f[:14]:14s}:{ln:4d} {ins / units:12.1f} {100 * smp / S:8.1f}% {n:5d}  {text}")
