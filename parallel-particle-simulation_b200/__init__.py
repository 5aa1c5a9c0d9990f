"""parallel-particle-simulation_b200 -- B200 (sm_100a) implementation of the reference's per-timestep
particle hot path behind its own `common.h` interface.

The product is native: `csrc/` holds the CUDA kernels, the C ABI (`include/psim.h`), the C++ shim that
exports the reference's `init_simulation` / `simulate_one_step` and a C++ driver.  This Python package is
only the test / bench harness binding (ctypes) -- it contains no compute and no CPU fallback.

The directory name contains '-', so import it through `__graft_entry__.load_package()` (or importlib).
"""
from .host.binding import (  # noqa: F401
    ENGINE_AUTO, ENGINE_CELLSORT, ENGINE_TILED, ENGINE_KSTEP, STEP_ACCEL_ALL, STEP_ACCEL_NONE, STEP_DEFAULT,
    PsimError, Simulation, bin_count, build_native, init_particles, init_simulation, lib, lib_path,
    simulate_one_step, box_size, generate_particles_device, comm_unique_id, slab_rows, DECLARED_SYMBOLS,
)
