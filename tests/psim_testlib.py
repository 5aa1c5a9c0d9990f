"""Shared test plumbing: ctypes bindings to the CHECKERS (oracle/ and oracle/_ref/).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg import this module.
The product (parallel-particle-simulation_b200/) never does.

* ``Oracle``      -- oracle/libpsim_oracle.so, the C restatement of reference part1/serial.cpp.
* ``RefKernel``   -- oracle/_ref/libref_{serial,naive,openmp}.so: the UNMODIFIED reference
                     kernels compiled by oracle/Makefile (present only if `make ref` was run in
                     the build container; the files travel to the GPU box with gpurun).
* ``ref_init_particles`` -- the reference driver's own generator (part1/main.cpp:31-59).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import shutil
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

DENSITY = 0.0005
CUTOFF = 0.01
BIN_SIZE = 0.01
DT = 0.0005
MASS = 0.01

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def box_size(n: int) -> float:
    """reference part1/main.cpp:113"""
    return math.sqrt(DENSITY * n)


def as_parts(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float64)
    assert a.ndim == 2 and a.shape[1] == 6
    return a


def _ptr(a: np.ndarray, typ=_dp):
    return a.ctypes.data_as(typ)


def ensure_oracle() -> str:
    so = os.path.join(ORACLE_DIR, "libpsim_oracle.so")
    src = os.path.join(ORACLE_DIR, "psim_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])
    return so


class Oracle:
    """ctypes view of oracle/psim_oracle.c (particles are (N,6) float64 arrays: x y vx vy ax ay)."""

    def __init__(self):
        self.lib = C.CDLL(ensure_oracle())
        L = self.lib
        L.orc_bin_count.restype = C.c_int
        L.orc_bin_count.argtypes = [C.c_double]
        L.orc_cell_ids.argtypes = [_dp, C.c_int, C.c_int, _ip]
        L.orc_cell_counts.argtypes = [_dp, C.c_int, C.c_int, _ip]
        L.orc_cell_lists.argtypes = [_dp, C.c_int, C.c_int, _ip, _ip]
        L.orc_compute_forces.argtypes = [_dp, C.c_int, C.c_double]
        L.orc_move.argtypes = [_dp, C.c_int, C.c_double]
        L.orc_simulate_one_step.argtypes = [_dp, C.c_int, C.c_double]
        L.orc_simulate_steps.argtypes = [_dp, C.c_int, C.c_double, C.c_int]
        L.orc_stats.argtypes = [_dp, C.c_int, C.c_double, _dp]

    def bin_count(self, size: float) -> int:
        return int(self.lib.orc_bin_count(size))

    def cell_ids(self, parts, size) -> np.ndarray:
        parts = as_parts(parts)
        out = np.empty(len(parts), dtype=np.int32)
        self.lib.orc_cell_ids(_ptr(parts), len(parts), self.bin_count(size), _ptr(out, _ip))
        return out

    def cell_counts(self, parts, size) -> np.ndarray:
        parts = as_parts(parts)
        nb = self.bin_count(size)
        out = np.empty(nb * nb, dtype=np.int32)
        self.lib.orc_cell_counts(_ptr(parts), len(parts), nb, _ptr(out, _ip))
        return out

    def cell_lists(self, parts, size):
        parts = as_parts(parts)
        nb = self.bin_count(size)
        start = np.empty(nb * nb + 1, dtype=np.int32)
        member = np.empty(max(len(parts), 1), dtype=np.int32)
        self.lib.orc_cell_lists(_ptr(parts), len(parts), nb, _ptr(start, _ip), _ptr(member, _ip))
        return start, member[: len(parts)]

    def compute_forces(self, parts, size) -> np.ndarray:
        """in place; returns parts"""
        assert parts.flags.c_contiguous and parts.dtype == np.float64
        self.lib.orc_compute_forces(_ptr(parts), len(parts), size)
        return parts

    def step(self, parts, size, steps: int = 1) -> np.ndarray:
        assert parts.flags.c_contiguous and parts.dtype == np.float64
        self.lib.orc_simulate_steps(_ptr(parts), len(parts), size, steps)
        return parts

    def stats(self, parts, size) -> dict:
        parts = as_parts(parts)
        out = np.zeros(8)
        self.lib.orc_stats(_ptr(parts), len(parts), size, _ptr(out))
        return dict(dmin=out[0], davg=out[1], pairs=int(out[2]), touched=int(out[3]), ke=out[4],
                    vmax=out[5], max_neighbours=int(out[6]))


def have_ref(name: str = "libref_serial.so") -> bool:
    return os.path.exists(os.path.join(REF_DIR, name))


class RefKernel:
    """One private instance of an unmodified reference kernel (.so copied so that every instance
    gets its own globals; dlopen'ed RTLD_LOCAL so that its init_simulation/simulate_one_step do
    not clash with the product's identically named symbols)."""

    INIT = "_Z15init_simulationP10particle_tid"
    STEP = "_Z17simulate_one_stepP10particle_tid"

    def __init__(self, flavour: str = "serial"):
        src = os.path.join(REF_DIR, f"libref_{flavour}.so")
        if not os.path.exists(src):
            raise FileNotFoundError(src)
        self._tmp = tempfile.NamedTemporaryFile(suffix=f"_{flavour}.so", delete=False)
        self._tmp.close()
        shutil.copyfile(src, self._tmp.name)
        self.lib = C.CDLL(self._tmp.name, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        os.unlink(self._tmp.name)
        self._init = getattr(self.lib, self.INIT)
        self._step = getattr(self.lib, self.STEP)
        self._init.argtypes = [_dp, C.c_int, C.c_double]
        self._step.argtypes = [_dp, C.c_int, C.c_double]
        self._init.restype = None
        self._step.restype = None
        self.flavour = flavour
        self.parts = None
        self.size = None

    def init(self, parts: np.ndarray, size: float):
        """parts is adopted: the reference keeps raw pointers into it (serial.cpp:86)."""
        assert parts.flags.c_contiguous and parts.dtype == np.float64 and parts.shape[1] == 6
        self.parts, self.size = parts, size
        self._init(_ptr(parts), len(parts), size)
        return self

    def step(self, steps: int = 1):
        for _ in range(steps):
            self._step(_ptr(self.parts), len(self.parts), self.size)
        return self.parts

    # --- accessors from oracle/ref_probe.cpp (serial / openmp flavours only) ---
    def bin_count(self) -> int:
        self.lib.ref_bin_count.restype = C.c_int
        return int(self.lib.ref_bin_count())

    def cell_counts(self) -> np.ndarray:
        nb = self.bin_count()
        out = np.empty(nb * nb, dtype=np.int32)
        self.lib.ref_cell_counts(_ptr(out, _ip))
        return out

    def cell_lists(self):
        nb = self.bin_count()
        start = np.empty(nb * nb + 1, dtype=np.int32)
        member = np.empty(max(len(self.parts), 1), dtype=np.int32)
        self.lib.ref_cell_lists(_ptr(self.parts), _ptr(start, _ip), _ptr(member, _ip))
        return start, member[: len(self.parts)]


_ref_driver = None


def ref_init_particles(n: int, seed: int) -> np.ndarray:
    """The reference driver's generator (part1/main.cpp:31-59); ax, ay zeroed (the driver leaves
    them uninitialised)."""
    global _ref_driver
    if _ref_driver is None:
        _ref_driver = C.CDLL(os.path.join(REF_DIR, "libref_driver.so"), mode=os.RTLD_LOCAL | os.RTLD_LAZY)
    f = getattr(_ref_driver, "_Z14init_particlesP10particle_tidi")
    f.argtypes = [_dp, C.c_int, C.c_double, C.c_int]
    f.restype = None
    parts = np.zeros((n, 6), dtype=np.float64)
    f(_ptr(parts), n, box_size(n), seed)
    parts[:, 4:] = 0.0
    return parts


def load_golden(name: str) -> np.ndarray:
    return np.load(os.path.join(GOLDEN_DIR, name))


def rel_err(a: np.ndarray, b: np.ndarray, floor: float = 1.0) -> float:
    """max |a-b| / max(|b|, floor) -- the tolerance form SURVEY.md section 8c states."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))
