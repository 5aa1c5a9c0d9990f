#!/usr/bin/env python
"""The reference's part2 (multi-process) flavour of the plugin interface, driven the way part2/main.cpp drives it:

    [python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1] tests/mpi_flavour_check.py [n] [steps]

Every rank loads libpsim_mpi_shim.so and calls the reference's own (C++-mangled) entry points
    init_simulation / simulate_one_step / gather_for_save (particle_t*, num_parts, size, rank, num_procs)
on 56-byte records {id, x, y, vx, vy, ax, ay} (part2/common.h:17-32); the torch launcher only provides RANK / WORLD_SIZE --
no MPI, no torch.distributed.  Every `savefreq` steps gather_for_save must leave the complete, id-ordered state on rank 0
(part2/main.cpp:158-166); rank 0 compares it BIT FOR BIT with the oracle.  Prints `MPI_FLAVOUR_CHECK ok ...`."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

REC = np.dtype([("id", "<u8"), ("x", "<f8"), ("y", "<f8"), ("vx", "<f8"), ("vy", "<f8"), ("ax", "<f8"), ("ay", "<f8")])


def main():
    import __graft_entry__ as g
    from psim_testlib import Oracle

    pkg = g.load_package()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    pkg.lib()   # libpsim.so first (RTLD_GLOBAL), then the shim that is linked against it
    shim = C.CDLL(os.path.join(os.path.dirname(pkg.lib_path()), "libpsim_mpi_shim.so"))
    sig = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int]
    init = getattr(shim, "_Z15init_simulationP10particle_tidii")
    step = getattr(shim, "_Z17simulate_one_stepP10particle_tidii")
    gather = getattr(shim, "_Z15gather_for_saveP10particle_tidii")
    for f in (init, step, gather):
        f.argtypes, f.restype = sig, None

    size = pkg.box_size(n)
    state = pkg.init_particles(n, 7)
    orc = Oracle()
    orc.step(state, size, 40)   # warmed state, identical on every rank (the reference broadcasts it, part2/main.cpp:150)
    parts = np.zeros(n, dtype=REC)
    parts["id"] = np.arange(1, n + 1)
    for k, name in enumerate(("x", "y", "vx", "vy")):
        parts[name] = state[:, k]
    want = state.copy()
    ok, checked = True, 0
    init(parts.ctypes.data, n, size, rank, world)
    for s in range(steps):
        step(parts.ctypes.data, n, size, rank, world)
        if rank == 0:
            orc.step(want, size, 1)
        if s % 10 == 0 or s == steps - 1:
            gather(parts.ctypes.data, n, size, rank, world)
            if rank == 0:
                got = np.stack([parts[k] for k in ("x", "y", "vx", "vy", "ax", "ay")], axis=1)
                same = bool(np.array_equal(got, want)) and bool(np.array_equal(parts["id"], np.arange(1, n + 1)))
                ok = ok and same
                checked += 1
    if rank == 0:
        print(f"MPI_FLAVOUR_CHECK {'ok' if ok else 'FAIL'} ranks={world} n={n} steps={steps} gathers_checked={checked}", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
