"""CPU tests (-m "not gpu"): the oracle against the reference fixtures and (when prebuilt) the live
reference kernels; the C ABI library loads and exports every declared symbol; host-side logic."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

from psim_testlib import GOLDEN_DIR, ROOT, RefKernel, box_size, have_ref, load_golden, ref_init_particles, rel_err


@pytest.mark.parametrize("fixture,marks", [("ref_n1000_s1.npz", [0, 1, 50, 51, 100, 300, 1000]),
                                           ("ref_n3000_s7.npz", [0, 100, 101, 200])])
def test_oracle_reproduces_reference_fixtures_bitwise(oracle, fixture, marks):
    g = load_golden(fixture)
    size = float(g["size"])
    state = g["step0"].copy()
    done = 0
    for m in marks:
        oracle.step(state, size, m - done)
        done = m
        assert np.array_equal(state, g[f"step{m}"]), f"step {m}"
        if f"cellid{m}" in g:
            assert np.array_equal(oracle.cell_ids(state, size), g[f"cellid{m}"])
            assert np.array_equal(oracle.cell_counts(state, size), g[f"cellcount{m}"])
    assert oracle.bin_count(size) == int(g["bincnt"])


def test_oracle_statistics_in_expected_band(oracle):
    g = load_golden("ref_n1000_s1.npz")
    st = oracle.stats(g["step1000"], float(g["size"]))
    assert 0.4 < st["dmin"] < 1.0 and 0.9 < st["davg"] < 1.0
    assert 0.02 < st["touched"] / 1000 < 0.25
    ke0 = oracle.stats(g["step0"], float(g["size"]))["ke"]
    assert abs(st["ke"] - ke0) / ke0 < 0.05  # kinetic energy drifts < 2 % over 1000 steps


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("flavour", ["serial", "naive", "openmp"])
def test_oracle_against_live_reference(oracle, flavour):
    n = 4000 if flavour != "naive" else 1500
    size = box_size(n)
    a = ref_init_particles(n, 9)
    b = a.copy()
    ref = RefKernel(flavour).init(a, size)
    for _ in range(4):
        ref.step(50)
        oracle.step(b, size, 50)
        assert rel_err(b[:, :4], a[:, :4]) <= 1e-12
        if flavour == "naive":
            # all-pairs sums run in index order, not cell order: a particle that had three in-range
            # neighbours may differ in the last bit
            assert int((a != b).any(axis=1).sum()) <= 3
        else:
            assert np.array_equal(a, b)
        if flavour != "naive":
            assert np.array_equal(ref.cell_counts(), oracle.cell_counts(b, size))
            rs, rm = ref.cell_lists()
            os_, om = oracle.cell_lists(b, size)
            assert np.array_equal(rs, os_)
            for c in np.nonzero(np.diff(rs) > 1)[0][:200]:
                assert sorted(rm[rs[c]:rs[c + 1]]) == sorted(om[os_[c]:os_[c + 1]])


def test_oracle_edge_cases(oracle):
    size = box_size(1000)
    # closer than min_r: r2 clamps to min_r^2 (serial.cpp:29) -> finite, huge repulsion
    p = np.array([[0.30, 0.30, 0, 0, 0, 0], [0.30000001, 0.30, 0, 0, 0, 0]], dtype=np.float64)
    oracle.compute_forces(p, size)
    assert np.isfinite(p).all() and p[0, 4] < 0 < p[1, 4] and abs(p[0, 4]) > 1e3
    # coincident particles feel nothing (coef * 0)
    p = np.array([[0.30, 0.30, 0, 0, 0, 0], [0.30, 0.30, 0, 0, 0, 0]], dtype=np.float64)
    oracle.compute_forces(p, size)
    assert (p[:, 4:] == 0).all()
    # reflection flips the velocity and keeps the particle inside
    p = np.array([[1e-5, size - 1e-5, -1.0, 1.0, 0, 0]], dtype=np.float64)
    oracle.step(p, size, 1)
    assert 0 <= p[0, 0] <= size and 0 <= p[0, 1] <= size and p[0, 2] == 1.0 and p[0, 3] == -1.0
    # cell index uses a true division: 0.07/0.01 and 0.29/0.01 are the classic x*100 mismatches
    ids = oracle.cell_ids(np.array([[0.07, 0.29, 0, 0, 0, 0], [0.58, 0.57, 0, 0, 0, 0]]), size)
    nb = oracle.bin_count(size)
    assert list(ids) == [int(np.floor(0.07 / 0.01)) * nb + int(np.floor(0.29 / 0.01)),
                         int(np.floor(0.58 / 0.01)) * nb + int(np.floor(0.57 / 0.01))]


# ---------------------------------------------------------------- the boundary (no compute without a GPU)
def test_library_exports_every_declared_symbol(pkg):
    L = pkg.lib()
    header = open(os.path.join(ROOT, "include", "psim.h")).read()
    declared = sorted(set(re.findall(r"\b(psim_[a-z_0-9]+)\s*\(", header)))
    assert declared == sorted(pkg.DECLARED_SYMBOLS)
    for s in declared:
        assert hasattr(L, s), s


def test_shim_exports_the_reference_mangled_names(pkg):
    shim = C.CDLL(os.path.join(os.path.dirname(pkg.lib_path()), "libpsim_shim.so"))
    assert hasattr(shim, "_Z15init_simulationP10particle_tid")
    assert hasattr(shim, "_Z17simulate_one_stepP10particle_tid")


def test_mpi_flavour_shim_exports_the_part2_mangled_names(pkg):
    """reference part2/common.h:30-32: rank-aware init_simulation / simulate_one_step and gather_for_save (C++ linkage)"""
    shim = C.CDLL(os.path.join(os.path.dirname(pkg.lib_path()), "libpsim_mpi_shim.so"))
    for name in ("_Z15init_simulationP10particle_tidii", "_Z17simulate_one_stepP10particle_tidii",
                 "_Z15gather_for_saveP10particle_tidii"):
        assert hasattr(shim, name), name


def test_no_cpu_fallback_without_a_device(pkg):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(pkg.PsimError) as ei:
        pkg.Simulation(np.zeros((4, 6)), 4, 0.5)
    assert ei.value.status == 2  # PSIM_ERR_NO_DEVICE


def test_generator_matches_reference_fixture(pkg):
    """psim_init_particles must reproduce the reference driver's init_particles bit for bit."""
    for fixture in ("ref_n1000_s1.npz", "ref_n3000_s7.npz"):
        g = load_golden(fixture)
        got = pkg.init_particles(int(g["n"]), int(g["seed"]))
        assert np.array_equal(got, g["step0"])


def test_save_frame_matches_reference_text_format(pkg, tmp_path):
    g = load_golden("ref_n1000_s1.npz")
    meta = json.load(open(os.path.join(GOLDEN_DIR, "ref_trajectory_n1000_s1.json")))
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    path = tmp_path / "frame.txt"
    f = libc.fopen(str(path).encode(), b"w")
    xy = np.ascontiguousarray(g["step1"][:, :2])
    assert pkg.lib().psim_save_frame(f, xy.ctypes.data, 1000, float(g["size"]), 1) == 0
    libc.fclose(f)
    lines = path.read_text().split("\n")
    assert lines[:4] == meta["head"]  # "1000 0.707107" + the first three particles of frame 0 (= state after step 1)
    assert len(lines) == 1000 + 3


def test_save_frame_formatter_equals_printf_g(pkg, tmp_path):
    """psim_save_frame formats with std::to_chars(general, 6); the reference streams doubles with the default ostream
    precision, i.e. printf("%g").  Same characters on random and on awkward values."""
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    rng = np.random.default_rng(1)
    vals = np.concatenate([rng.random(50000) * 300, 10.0 ** rng.uniform(-12, 3, 50000),
                           [0.0, 1e-5, 9.99999e-5, 0.0001, 0.000123456789, 999999.5, 1e6, 123456.5, 0.1, 100.0, 99.99995,
                            0.99999949, 5e-324, 1.0, 282.84271247461902, 0.70710678118654757]])
    xy = np.ascontiguousarray(vals.reshape(-1, 2))
    path = tmp_path / "f.txt"
    f = libc.fopen(str(path).encode(), b"w")
    assert pkg.lib().psim_save_frame(f, xy.ctypes.data, len(xy), 1.0, 0) == 0
    libc.fclose(f)
    got = path.read_text().split("\n")[:-2]
    assert got == ["%g %g" % (a, b) for a, b in xy]
