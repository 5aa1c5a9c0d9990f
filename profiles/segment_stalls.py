#!/usr/bin/env python
"""Per-segment stall mix of one kernel in an ncu report (segments end at barriers / mbarrier waits / bulk copies).
Usage: segment_stalls.py report.ncu-rep units_per_launch"""
import csv, subprocess, sys, io
rep, units = sys.argv[1], float(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]; data = [r for r in rows[2:] if len(r) >= len(h)]
si = h.index("# Samples"); ii = h.index("Instructions Executed"); src = h.index("Source")
stall = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
S = sum(int(r[si]) for r in data)
seg_i = seg_s = 0; seg = {}
for k, r in enumerate(data):
    seg_i += int(r[ii]); seg_s += int(r[si]); t = r[src]
    for i in stall: seg[h[i]] = seg.get(h[i], 0) + int(r[i] or 0)
    if 'BAR.SYNC' in t or 'SYNCS.PHASECHK' in t or 'EXIT' in t:
        if seg_s > 0.01 * S:
            top = sorted(seg.items(), key=lambda kv: -kv[1])[:5]
            print(f"{k:5d} instr/unit {seg_i / units:9.1f} samples {100 * seg_s / S:5.1f}%  " + " ".join(f"{a[6:]}={100 * b / S:.1f}" for a, b in top) + f"   | {t[:40]}")
        seg_i = seg_s = 0; seg = {}
