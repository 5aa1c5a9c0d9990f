#!/usr/bin/env python
"""Long-run soak: step in chunks, synchronise, print the engine's high-water marks (debug aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as g
pkg = g.load_package()
n = int(sys.argv[1]); total = int(sys.argv[2]); chunk = int(sys.argv[3]); tile = int(sys.argv[4]) if len(sys.argv) > 4 else 0
flag = int(sys.argv[5]) if len(sys.argv) > 5 else pkg.STEP_DEFAULT
size = pkg.box_size(n)
parts = pkg.init_particles(n, 42)
sim = pkg.Simulation(parts, n, size, engine=pkg.ENGINE_TILED, tile_cells=tile)
done = 0
while done < total:
    sim.step(chunk, flag).sync()
    done += chunk
    i = sim.info()
    print(done, {k: i[k] for k in ("hw_leavers", "hw_halo_list", "hw_tile_population", "hw_apron", "reserved_hw_pairs")}, flush=True)
st = sim.stats()
print("stats", st)
