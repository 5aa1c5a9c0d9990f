#!/usr/bin/env python
"""Does an idle gap before psim_create slow its host->device copy down (PCIe link / GPU power state)?  Two pinned buffers, so
the passes can run back to back: create + 20 steps + read-back, after gaps of 0 / 3 s, with and without an nvidia-smi poller."""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import __graft_entry__ as g
pkg = g.load_package()
n, steps = 20_000_000, 20
size = pkg.box_size(n)
src = torch.empty((n, 6), dtype=torch.float64, pin_memory=True)
dst = torch.empty((n, 6), dtype=torch.float64, pin_memory=True)
pkg.init_particles(n, 42, size, out=src.numpy())
torch.cuda.synchronize()
def one(tag):
    t0 = time.perf_counter(); sim = pkg.Simulation(src, n, size); t1 = time.perf_counter()
    sim.step(steps).sync(); t2 = time.perf_counter()
    sim.read_particles(dst); t3 = time.perf_counter()
    sim.close()
    print(f"{tag:28s} create {1e3*(t1-t0):7.1f} ms  steps {1e3*(t2-t1):6.1f} ms  read {1e3*(t3-t2):6.1f} ms  total {1e3*(t3-t0):7.1f} ms", flush=True)
one("first")
one("back to back"); one("back to back")
time.sleep(3); one("after 3 s idle")
time.sleep(3); one("after 3 s idle")
one("back to back")
t = time.time()
while time.time() - t < 3: sum(range(100000))   # busy CPU, idle GPU
one("after 3 s of CPU work")
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,pcie.link.gen.current,pcie.link.width.current", "--format=csv,noheader", "-lms", "100"], stdout=subprocess.PIPE, text=True)
time.sleep(1.0); p.terminate(); out = p.stdout.read().strip().splitlines(); print("smi idle:", out[-1] if out else None)
one("right after nvidia-smi poll")
time.sleep(3); one("3 s after nvidia-smi")
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,pcie.link.gen.current,pcie.link.width.current", "--format=csv,noheader", "-lms", "20"], stdout=subprocess.PIPE, text=True)
time.sleep(0.5); one("while nvidia-smi polls"); p.terminate(); out = p.stdout.read().strip().splitlines(); print("smi during:", out[-3:])
