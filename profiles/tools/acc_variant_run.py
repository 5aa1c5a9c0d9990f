#!/usr/bin/env python
"""Steady-state launches of the acceleration-storing kernel variant, for ncu: 300 untimed steps (ACCEL_NONE, 100 launches of
kstep_kernel<64,4,false,false>), then five 3-step batches with PSIM_STEP_DEFAULT, each of which is ONE launch of
kstep_kernel<64,4,true,false> (ax, ay materialised for the last step of the batch).  No torch: ctypes binding only.

    ncu --set full -k regex:kstep_kernel --launch-skip 103 --launch-count 1 -o acc python profiles/tools/acc_variant_run.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as g
pkg = g.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
size = pkg.box_size(n)
parts = pkg.init_particles(n, 42)
sim = pkg.Simulation(parts, n, size)
sim.step(300, pkg.STEP_ACCEL_NONE).sync()
for _ in range(5):
    sim.step(3, pkg.STEP_DEFAULT).sync()
i = sim.info()
print("launches", i["kernel_launches"], "steps", i["steps_done"], "engine", i["engine"], "tile", i["tile_cells"])
sim.close()
