#!/usr/bin/env python
"""Small deterministic run for compute-sanitizer: a warmed state, a few steps on the tiled engine, compared with the oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
from psim_testlib import Oracle
pkg = g.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
tile = int(sys.argv[3]) if len(sys.argv) > 3 else 16
size = pkg.box_size(n)
parts = pkg.init_particles(n, 5)
orc = Oracle(); orc.step(parts, size, 40)
want = parts.copy(); orc.step(want, size, steps)
sim = pkg.Simulation(parts, n, size, engine=pkg.ENGINE_TILED, tile_cells=tile)
got = sim.step(steps).sync().read_particles()
print("bit_identical", bool(np.array_equal(got, want)), sim.info()["kernel_launches"])
sim.close()
