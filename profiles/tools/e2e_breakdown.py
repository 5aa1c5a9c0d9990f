#!/usr/bin/env python
"""Where the end-to-end second goes: psim_create (H2D + tiling), the steps, psim_read_particles (gather + D2H)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import __graft_entry__ as g
pkg = g.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
size = pkg.box_size(n)
host = torch.empty((n, 6), dtype=torch.float64, pin_memory=True)
pkg.init_particles(n, 42, size, out=host.numpy())
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter(); sim = pkg.Simulation(host, n, size); t1 = time.perf_counter()
    sim.step(steps).sync(); t2 = time.perf_counter()
    sim.read_particles(host); t3 = time.perf_counter()
    sim.close(); t4 = time.perf_counter()
    print(f"create {1e3*(t1-t0):7.1f} ms  steps {1e3*(t2-t1):7.1f} ms  read {1e3*(t3-t2):7.1f} ms  close {1e3*(t4-t3):6.1f} ms  total {1e3*(t3-t0):7.1f} ms", flush=True)
    pkg.init_particles(n, 42, size, out=host.numpy())
