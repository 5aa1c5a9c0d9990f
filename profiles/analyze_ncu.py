#!/usr/bin/env python
"""Summarise an ncu report of tile_step_kernel: stall mix, warp-instructions per tile per phase
(segments between barriers), hottest SASS instructions.  Usage: analyze_ncu.py report.ncu-rep ntiles"""
import csv, subprocess, sys, io
rep, tiles = sys.argv[1], float(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]; data = [r for r in rows[2:] if len(r) >= len(h)]
si = h.index("# Samples"); ii = h.index("Instructions Executed"); src = h.index("Source")
stall = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
S = sum(int(r[si]) for r in data)
tot = {}
for r in data:
    for i in stall: tot[h[i]] = tot.get(h[i], 0) + int(r[i] or 0)
print("stall mix %:", {k: round(100 * v / S, 1) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]})
seg_i = seg_s = 0
print("segment-end  warp-instr/tile  samples%")
for k, r in enumerate(data):
    seg_i += int(r[ii]); seg_s += int(r[si]); t = r[src]
    if 'BAR.SYNC' in t or 'SYNCS' in t or 'UBLKCP' in t or 'EXIT' in t:
        if seg_i / tiles > 5 or seg_s > 0.01 * S:
            print(f"{k:5d} {seg_i / tiles:9.1f} {100 * seg_s / S:6.1f}%  {t[:60]}")
        seg_i = seg_s = 0
print("total warp-instr/tile", round(sum(int(r[ii]) for r in data) / tiles, 1), " SASS instrs", len(data))
print("hottest:")
for k, r in enumerate(data):
    if int(r[si]) > 0.012 * S: print(f"{k:5d} {int(r[si]):7d} {r[ii]:>10s}  {r[src][:70]}")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw))); hh = rr[0]; vv = rr[-1]
for key in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "smsp__inst_executed_op_shared_atom.sum", "lts__t_sector_hit_rate.pct"):
    if key in hh: print(key, rr[1][hh.index(key)], vv[hh.index(key)])
