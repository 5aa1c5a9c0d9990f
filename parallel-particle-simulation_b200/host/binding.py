"""ctypes binding of libpsim's C ABI (include/psim.h) plus a thin mirror of the reference interface.

Mirror of the reference's plugin interface (reference part1/common.h:24-25):

    init_simulation(parts, num_parts, size)     -> builds the device state from an (N,6) float64 array
    simulate_one_step(parts, num_parts, size)   -> one step; `parts` is updated in place like the
                                                   reference updates its caller's array

Everything that computes goes through the CUDA library; if it is missing or no GPU is visible the
calls raise -- there is deliberately no CPU path here.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(_HERE), "csrc")
BUILD = os.path.join(CSRC, "build")

ENGINE_AUTO, ENGINE_CELLSORT, ENGINE_TILED, ENGINE_KSTEP = 0, 1, 2, 3
STEP_DEFAULT, STEP_ACCEL_ALL, STEP_ACCEL_NONE = 0, 1, 2

# every symbol include/psim.h declares (tests check that the library exports each of them)
DECLARED_SYMBOLS = [
    "psim_error_string", "psim_last_error", "psim_device_init", "psim_host_register", "psim_host_unregister", "psim_config_default", "psim_bin_count", "psim_create",
    "psim_destroy", "psim_step", "psim_sync", "psim_read_particles", "psim_read_positions",
    "psim_read_cells", "psim_read_cell_lists", "psim_stats", "psim_info", "psim_init_particles",
    "psim_save_frame", "psim_comm_unique_id", "psim_comm_connect", "psim_slab_rows", "psim_state_hash", "psim_gather", "psim_device_count", "psim_read_positions_begin", "psim_read_positions_end", "psim_generate_particles_device",
]


class PsimError(RuntimeError):
    def __init__(self, status: int, what: str, detail: str):
        super().__init__(f"{what}: status {status} ({detail})")
        self.status = status


class Config(C.Structure):
    _fields_ = [("engine", C.c_int), ("device", C.c_int), ("stream", C.c_void_p), ("tile_cells", C.c_int),
                ("steps_per_launch", C.c_int), ("rank", C.c_int), ("nranks", C.c_int), ("reserved", C.c_int * 8)]


class Stats(C.Structure):
    _fields_ = [("dmin", C.c_double), ("davg", C.c_double), ("kinetic_energy", C.c_double), ("vmax", C.c_double),
                ("pairs", C.c_longlong), ("touched", C.c_longlong), ("max_neighbours", C.c_int),
                ("max_cell_count", C.c_int)]


class Info(C.Structure):
    _fields_ = [("engine", C.c_int), ("bin_count", C.c_int), ("tile_cells", C.c_int), ("tiles_per_side", C.c_int),
                ("tile_capacity", C.c_int), ("device", C.c_int), ("num_parts", C.c_int), ("rank", C.c_int),
                ("nranks", C.c_int), ("row_begin", C.c_int), ("row_end", C.c_int), ("steps_done", C.c_longlong),
                ("kernel_launches", C.c_longlong), ("device_bytes", C.c_longlong), ("hw_leavers", C.c_int),
                ("hw_halo_list", C.c_int), ("hw_tile_population", C.c_int), ("hw_apron", C.c_int),
                ("outbox_capacity", C.c_int), ("halo_list_capacity", C.c_int), ("reserved_hw_pairs", C.c_int),
                ("halo_cells", C.c_int), ("steps_per_launch", C.c_int), ("region_capacity", C.c_int),
                ("recoveries", C.c_int), ("input_on_device", C.c_int), ("engine_switches", C.c_int)]


def lib_path() -> str:
    # PSIM_LIB selects another build of the same library (e.g. the phase-timer profiling build)
    return os.environ.get("PSIM_LIB") or os.path.join(BUILD, "libpsim.so")


def build_native(force: bool = False) -> str:
    """Compile csrc/ for sm_100a (nvcc cross-compiles without a GPU)."""
    if force:
        subprocess.check_call(["make", "-s", "-C", CSRC, "clean"])
    subprocess.check_call(["make", "-s", "-j8", "-C", CSRC])
    return lib_path()


_lib = None


def lib() -> C.CDLL:
    """The C ABI library.  Raises if it has not been built -- no fallback exists."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                "(libpsim is the only compute path; there is no CPU fallback)")
    L = C.CDLL(path, mode=os.RTLD_GLOBAL)
    L.psim_error_string.restype = C.c_char_p
    L.psim_error_string.argtypes = [C.c_int]
    L.psim_last_error.restype = C.c_char_p
    L.psim_config_default.argtypes = [C.POINTER(Config)]
    L.psim_config_default.restype = None
    L.psim_bin_count.argtypes = [C.c_double]
    L.psim_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(Config), C.c_void_p, C.c_int, C.c_double]
    L.psim_destroy.argtypes = [C.c_void_p]
    L.psim_step.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.psim_sync.argtypes = [C.c_void_p]
    L.psim_read_particles.argtypes = [C.c_void_p, C.c_void_p]
    L.psim_read_positions.argtypes = [C.c_void_p, C.c_void_p]
    L.psim_read_cells.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.psim_read_cell_lists.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.psim_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    L.psim_info.argtypes = [C.c_void_p, C.POINTER(Info)]
    L.psim_read_positions_begin.argtypes = [C.c_void_p, C.c_void_p]
    L.psim_read_positions_end.argtypes = [C.c_void_p]
    L.psim_gather.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.psim_state_hash.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong), C.POINTER(C.c_longlong)]
    L.psim_init_particles.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int]
    L.psim_generate_particles_device.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_void_p]
    L.psim_save_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int]
    L.psim_comm_unique_id.argtypes = [C.c_void_p]
    L.psim_comm_connect.argtypes = [C.c_void_p, C.c_void_p]
    L.psim_slab_rows.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    _lib = L
    return L


def _check(status: int, what: str):
    if status != 0:
        L = lib()
        raise PsimError(status, what, f"{L.psim_error_string(status).decode()}: {L.psim_last_error().decode()}")


def box_size(n: int) -> float:
    """reference part1/main.cpp:113"""
    return math.sqrt(0.0005 * n)


def bin_count(size: float) -> int:
    return int(lib().psim_bin_count(size))


def init_particles(n: int, seed: int, size: float | None = None, out: np.ndarray | None = None) -> np.ndarray:
    """The reference driver's generator (part1/main.cpp:31-59) -> (N,6) float64 array."""
    size = box_size(n) if size is None else size
    parts = np.zeros((n, 6), dtype=np.float64) if out is None else out
    assert parts.flags.c_contiguous and parts.dtype == np.float64 and parts.shape == (n, 6)
    _check(lib().psim_init_particles(parts.ctypes.data, n, size, seed), "psim_init_particles")
    return parts


def generate_particles_device(out, n: int, seed: int, size: float | None = None, stream: int | None = None):
    """alternative seeding mode: the reference's construction generated in parallel into a DEVICE array (torch cuda tensor
    of shape (n, 6), float64)"""
    size = box_size(n) if size is None else size
    _check(lib().psim_generate_particles_device(_address(out), n, size, seed, stream), "psim_generate_particles_device")
    return out


def _address(buf) -> int:
    """host numpy array, torch tensor (host or cuda) or raw integer address"""
    if isinstance(buf, int):
        return buf
    if isinstance(buf, np.ndarray):
        assert buf.flags.c_contiguous
        return buf.ctypes.data
    if hasattr(buf, "data_ptr"):
        assert buf.is_contiguous()
        return buf.data_ptr()
    raise TypeError(type(buf))


class Simulation:
    """One simulation handle (psim_create ... psim_destroy)."""

    def __init__(self, parts, num_parts: int | None = None, size: float | None = None, *, engine: int = ENGINE_AUTO,
                 device: int = -1, stream: int | None = None, tile_cells: int = 0, rank: int = 0, nranks: int = 1,
                 steps_per_launch: int = 0):
        L = lib()
        n = int(num_parts if num_parts is not None else len(parts))
        self.n = n
        self.size = float(box_size(n) if size is None else size)
        cfg = Config()
        L.psim_config_default(C.byref(cfg))
        cfg.engine, cfg.device, cfg.tile_cells, cfg.rank, cfg.nranks = engine, device, tile_cells, rank, nranks
        cfg.steps_per_launch = steps_per_launch
        cfg.stream = stream
        self._parts_ref = parts   # slabs upload cooperatively inside comm_connect: keep the caller's array alive until then
        self._h = C.c_void_p()
        _check(L.psim_create(C.byref(self._h), C.byref(cfg), _address(parts) if n else None, n, self.size), "psim_create")

    # reference part1/serial.cpp:119-131 x nsteps
    def step(self, nsteps: int = 1, flags: int = STEP_DEFAULT):
        _check(lib().psim_step(self._h, nsteps, flags), "psim_step")
        return self

    def sync(self):
        _check(lib().psim_sync(self._h), "psim_sync")
        return self

    def read_particles(self, out=None):
        if out is None:
            out = np.zeros((self.n, 6), dtype=np.float64)
        _check(lib().psim_read_particles(self._h, _address(out)), "psim_read_particles")
        return out

    def read_positions(self, out=None):
        if out is None:
            out = np.zeros((self.n, 2), dtype=np.float64)
        _check(lib().psim_read_positions(self._h, _address(out)), "psim_read_positions")
        return out

    def read_positions_begin(self, out):
        _check(lib().psim_read_positions_begin(self._h, _address(out)), "psim_read_positions_begin")
        return out

    def read_positions_end(self):
        _check(lib().psim_read_positions_end(self._h), "psim_read_positions_end")
        return self

    def read_cells(self, want_counts: bool = True):
        nb = bin_count(self.size)
        ids = np.empty(self.n, dtype=np.int32)
        counts = np.empty(nb * nb, dtype=np.int32) if want_counts else None
        _check(lib().psim_read_cells(self._h, ids.ctypes.data, counts.ctypes.data if want_counts else None),
               "psim_read_cells")
        return ids, counts

    def read_cell_lists(self):
        nb = bin_count(self.size)
        start = np.empty(nb * nb + 1, dtype=np.int32)
        members = np.empty(max(self.n, 1), dtype=np.int32)
        _check(lib().psim_read_cell_lists(self._h, start.ctypes.data, members.ctypes.data), "psim_read_cell_lists")
        return start, members[: start[-1]]

    def stats(self) -> dict:
        s = Stats()
        _check(lib().psim_stats(self._h, C.byref(s)), "psim_stats")
        return {k: getattr(s, k) for k, _ in Stats._fields_}

    def gather(self, out=None, root: int = 0):
        """collective: rank `root` gets all particles in original order (reference part2 gather_for_save)"""
        _check(lib().psim_gather(self._h, _address(out) if out is not None else None, root), "psim_gather")
        return out

    def state_hash(self) -> tuple[int, int]:
        """(order-independent 64-bit fingerprint of the owned particles, number of owned particles)"""
        h, n = C.c_ulonglong(), C.c_longlong()
        _check(lib().psim_state_hash(self._h, C.byref(h), C.byref(n)), "psim_state_hash")
        return int(h.value), int(n.value)

    def info(self) -> dict:
        i = Info()
        _check(lib().psim_info(self._h, C.byref(i)), "psim_info")
        return {k: getattr(i, k) for k, _ in Info._fields_}

    def comm_connect(self, unique_id: bytes):
        buf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        _check(lib().psim_comm_connect(self._h, buf), "psim_comm_connect")
        self._parts_ref = None

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().psim_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def slab_rows(bin_count_: int, rank: int, nranks: int, tile_cells: int = 0) -> tuple[int, int]:
    """cell rows [begin, end) of slab `rank` (host-side arithmetic only; reference precedent part2/mpi.cpp:258-270)"""
    b, e = C.c_int(), C.c_int()
    _check(lib().psim_slab_rows(bin_count_, tile_cells, rank, nranks, C.byref(b), C.byref(e)), "psim_slab_rows")
    return b.value, e.value


def comm_unique_id() -> bytes:
    buf = (C.c_ubyte * 128)()
    _check(lib().psim_comm_unique_id(buf), "psim_comm_unique_id")
    return bytes(buf)


# ---- mirror of the reference's two entry points (part1/common.h:24-25) -------------------------------
_current: Simulation | None = None


def init_simulation(parts, num_parts: int, size: float, **kw) -> None:
    global _current
    if _current is not None:
        _current.close()
    _current = Simulation(parts, num_parts, size, **kw)


def simulate_one_step(parts, num_parts: int, size: float) -> None:
    """One step, then the caller's array is brought up to date (the literal contract: the reference
    mutates `parts` in place, part1/serial.cpp:127-130)."""
    if _current is None:
        raise PsimError(5, "simulate_one_step", "init_simulation has not been called")
    _current.step(1, STEP_DEFAULT)
    _current.read_particles(parts)
