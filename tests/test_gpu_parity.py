"""GPU parity tests: the CUDA path (through the C ABI) against the oracle, the committed reference
fixtures (tests/golden/, produced from the unmodified reference by gen_golden.py) and -- where the
prebuilt files travelled with the snapshot -- the reference kernels themselves (oracle/_ref/).

Tolerances (BASELINE.json north_star): cell ids and per-cell counts bit-exact; positions, velocities
and accelerations after one step within 1e-12 relative (|d| / max(|ref|, 1)); in practice the CUDA
arithmetic is non-contracted IEEE double in a canonical summation order, so the tests also assert
BIT equality wherever the reference's own summation order is reproducible (<= 2 in-range neighbours,
which is every particle of the fixtures)."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from psim_testlib import GOLDEN_DIR, REF_DIR, RefKernel, box_size, have_ref, load_golden, rel_err

pytestmark = pytest.mark.gpu

ENGINES = ["cellsort", "tiled16", "tiled32", "tiled64", "kstep16", "kstep32", "kstep48", "kstep64"]


def make_sim(pkg, parts, size, engine):
    if engine == "cellsort":
        return pkg.Simulation(parts, len(parts), size, engine=pkg.ENGINE_CELLSORT)
    if engine.startswith("kstep"):
        return pkg.Simulation(parts, len(parts), size, engine=pkg.ENGINE_KSTEP, tile_cells=int(engine[5:]))
    return pkg.Simulation(parts, len(parts), size, engine=pkg.ENGINE_TILED, tile_cells=int(engine[5:]))


# ---------------------------------------------------------------- binning
@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("fixture,step", [("ref_n1000_s1.npz", 0), ("ref_n1000_s1.npz", 100), ("ref_n1000_s1.npz", 1000),
                                          ("ref_n3000_s7.npz", 200)])
def test_cells_match_reference_bins_bit_exact(pkg, engine, fixture, step):
    g = load_golden(fixture)
    parts, size = g[f"step{step}"].copy(), float(g["size"])
    sim = make_sim(pkg, parts, size, engine)
    ids, counts = sim.read_cells()
    assert pkg.bin_count(size) == int(g["bincnt"])
    assert np.array_equal(ids, g[f"cellid{step}"])          # the reference's own Bins membership
    assert np.array_equal(counts, g[f"cellcount{step}"])    # the sizes of the reference's sets
    sim.close()


@pytest.mark.parametrize("engine", ["cellsort", "tiled32", "kstep32"])
def test_cell_lists_match_oracle(pkg, oracle, engine):
    n = 20000
    size = box_size(n)
    parts = pkg.init_particles(n, 5)
    oracle.step(parts, size, 40)
    sim = make_sim(pkg, parts, size, engine)
    start, members = sim.read_cell_lists()
    ostart, omembers = oracle.cell_lists(parts, size)
    assert np.array_equal(start, ostart)
    # same membership per cell; ours is ordered by original index, the oracle's by (x, y, index)
    order = np.lexsort((omembers, np.repeat(np.arange(len(ostart) - 1), np.diff(ostart))))
    assert np.array_equal(members, omembers[order])
    sim.close()


# ---------------------------------------------------------------- one step from a warmed state (a6-a9)
@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("fixture,step", [("ref_n1000_s1.npz", 50), ("ref_n3000_s7.npz", 100)])
def test_one_step_matches_reference(pkg, oracle, engine, fixture, step):
    g = load_golden(fixture)
    parts, size = g[f"step{step}"].copy(), float(g["size"])
    want = g[f"step{step + 1}"]
    sim = make_sim(pkg, parts, size, engine)
    got = sim.step(1).sync().read_particles()
    assert rel_err(got[:, :2], want[:, :2]) <= 1e-12     # positions
    assert rel_err(got[:, 2:4], want[:, 2:4]) <= 1e-12   # velocities
    assert rel_err(got[:, 4:], want[:, 4:]) <= 1e-12     # accelerations used by the step
    assert np.abs(want[:, 4:]).max() > 0                  # the fixture really has interactions
    assert np.array_equal(got, want), "expected bit equality with the reference for this fixture"
    sim.close()


# ---------------------------------------------------------------- trajectories
@pytest.mark.parametrize("engine", ENGINES)
def test_trajectory_1000_steps_bit_identical_to_reference(pkg, engine):
    """BASELINE configs[0]: -n 1000 -s 1, 1000 steps.  serial.cpp and the canonical order agree bit
    for bit at this size (no particle ever has three in-range neighbours)."""
    g = load_golden("ref_n1000_s1.npz")
    size = float(g["size"])
    sim = make_sim(pkg, g["step0"].copy(), size, engine)
    done = 0
    for mark in (1, 50, 100, 300, 1000):
        sim.step(mark - done)
        done = mark
        got = sim.sync().read_particles()
        assert np.array_equal(got[:, :4], g[f"step{mark}"][:, :4]), f"diverged by step {mark}"
        assert np.array_equal(got[:, 4:], g[f"step{mark}"][:, 4:])
    sim.close()


@pytest.mark.parametrize("engine", ENGINES)
def test_short_trajectory_vs_oracle_20k(pkg, oracle, engine):
    """100 steps at N = 20 000: stated tolerance 1e-12 relative on positions; the canonical summation
    order makes CUDA and oracle bit-identical, which is asserted too."""
    n = 20000
    size = box_size(n)
    parts = pkg.init_particles(n, 11)
    oracle.step(parts, size, 30)
    want = parts.copy()
    sim = make_sim(pkg, parts, size, engine)
    oracle.step(want, size, 100)
    got = sim.step(100).sync().read_particles()
    assert rel_err(got[:, :4], want[:, :4]) <= 1e-12
    assert np.array_equal(got, want)
    sim.close()


def test_engines_agree_bitwise_and_with_reference_statistics(pkg, oracle):
    n = 50000
    size = box_size(n)
    parts = pkg.init_particles(n, 42)
    sims = {e: make_sim(pkg, parts, size, e) for e in ENGINES}
    states = {e: s.step(300).sync().read_particles() for e, s in sims.items()}
    for e in ENGINES[1:]:
        assert np.array_equal(states[e], states["cellsort"]), e
    want = parts.copy()
    oracle.step(want, size, 300)
    assert np.array_equal(states["cellsort"], want)
    st, ost = sims["tiled32"].stats(), oracle.stats(want, size)
    assert st["pairs"] == ost["pairs"] and st["touched"] == ost["touched"]
    assert st["max_neighbours"] == ost["max_neighbours"]
    assert abs(st["dmin"] - ost["dmin"]) <= 1e-15 and abs(st["davg"] - ost["davg"]) <= 1e-12
    assert abs(st["kinetic_energy"] - ost["ke"]) <= 1e-9 * ost["ke"]
    assert abs(st["vmax"] - ost["vmax"]) <= 1e-12
    for s in sims.values():
        s.close()


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference at build time)")
def test_against_live_reference_kernel_100k(pkg):
    """BASELINE configs[1] size: 100 000 particles, 60 steps, against the unmodified serial.cpp."""
    n = 100000
    size = box_size(n)
    parts = pkg.init_particles(n, 42)
    ref_state = parts.copy()
    ref = RefKernel("serial").init(ref_state, size)
    sim = pkg.Simulation(parts, n, size)
    ref.step(60)
    got = sim.step(60).sync().read_particles()
    assert rel_err(got[:, :4], ref_state[:, :4]) <= 1e-12
    differing = int((got != ref_state).any(axis=1).sum())
    assert differing <= 5  # only particles with >= 3 in-range neighbours sharing a cell may differ in the last bit
    ids, counts = sim.read_cells()
    if differing == 0:
        assert np.array_equal(counts, ref.cell_counts())
    sim.close()


def test_1m_particles_100_steps_bit_identical_to_oracle(pkg, oracle):
    """BASELINE configs[2] size: 1 M particles, the first 100 steps from the reference's lattice start (seed 42, the
    reference's job scripts), default engine, against the oracle: stated tolerance 1e-12 relative, asserted bit-identical."""
    n = 1_000_000
    size = box_size(n)
    parts = pkg.init_particles(n, 42)
    want = parts.copy()
    sim = pkg.Simulation(parts, n, size)
    got = sim.step(100).sync().read_particles()
    oracle.step(want, size, 100)
    assert rel_err(got[:, :4], want[:, :4]) <= 1e-12
    assert np.array_equal(got, want)
    assert np.abs(want[:, 4:]).max() > 0
    st, ost = sim.stats(), oracle.stats(want, size)
    assert st["pairs"] == ost["pairs"] and abs(st["dmin"] - ost["dmin"]) <= 1e-15
    sim.close()


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference at build time)")
def test_20m_particles_vs_live_reference_serial(pkg):
    """BASELINE configs[3] -- the headline size: 20 M particles.  The GPU warms the lattice start up for 60 steps; from that
    state the UNMODIFIED reference part1/serial.cpp (oracle/_ref/libref_serial.so, in memory) and the default engine each
    advance 5 steps.  Tolerance 1e-12 relative on x, y, vx, vy; the only rows allowed to differ at all are particles with
    three or more in-range neighbours (the reference sums them in hash-set order)."""
    n = 20_000_000
    size = box_size(n)
    parts = pkg.init_particles(n, 42)
    sim = pkg.Simulation(parts, n, size)
    warmed = sim.step(60).sync().read_particles()
    ref_state = warmed.copy()
    ref = RefKernel("serial").init(ref_state, size)
    ref.step(5)
    got = sim.step(5).sync().read_particles()
    info = sim.info()
    sim.close()
    assert info["engine"] == pkg.ENGINE_KSTEP and info["engine_switches"] == 0
    assert rel_err(got[:, :4], ref_state[:, :4]) <= 1e-12
    assert rel_err(got[:, 4:], ref_state[:, 4:], floor=1e3) <= 1e-12   # accelerations of the last step (|a| up to ~5e3)
    differing = int((got != ref_state).any(axis=1).sum())
    assert differing <= 200, differing
    assert np.abs(ref_state[:, 4:]).max() > 0


# ---------------------------------------------------------------- edge cases
@pytest.mark.parametrize("engine", ["cellsort", "tiled16", "kstep16", "kstep64"])
def test_edge_cases(pkg, oracle, engine):
    size = box_size(1000)
    cases = {
        "single": np.array([[0.3, 0.3, 0.5, -0.25, 0, 0]]),
        "pair_in_range": np.array([[0.300, 0.300, 0, 0, 0, 0], [0.305, 0.302, 0, 0, 0, 0]]),
        "pair_closer_than_min_r": np.array([[0.30, 0.30, 0, 0, 0, 0], [0.30000001, 0.30, 0, 0, 0, 0]]),
        "coincident": np.array([[0.30, 0.30, 0.1, 0, 0, 0], [0.30, 0.30, -0.1, 0, 0, 0]]),
        "wall_bounce": np.array([[1e-5, size - 1e-5, -1.0, 1.0, 0, 0], [size - 1e-6, 2e-6, 3.0, -2.5, 0, 0]]),
        "four_body": np.array([[0.40, 0.40, 0, 0, 0, 0], [0.405, 0.40, 0, 0, 0, 0], [0.40, 0.406, 0, 0, 0, 0],
                               [0.396, 0.397, 0, 0, 0, 0], [0.4041, 0.4043, 0, 0, 0, 0]]),
        "cell_boundary": np.array([[0.07, 0.29, 0, 0, 0, 0], [0.58, 0.57, 0, 0, 0, 0], [0.0, 0.0, 0, 0, 0, 0]]),
    }
    for name, parts in cases.items():
        parts = np.ascontiguousarray(parts, dtype=np.float64)
        want = parts.copy()
        sim = make_sim(pkg, parts, size, engine)
        ids, counts = sim.read_cells()
        assert np.array_equal(ids, oracle.cell_ids(parts, size)), name
        assert np.array_equal(counts, oracle.cell_counts(parts, size)), name
        for _ in range(5):
            oracle.step(want, size, 1)
            got = sim.step(1).sync().read_particles()
            assert np.array_equal(got, want), name
        sim.close()


@pytest.mark.parametrize("engine", ["cellsort", "tiled16", "tiled32", "kstep16", "kstep64"])
def test_pair_list_overflow_takes_the_exact_path(pkg, oracle, engine):
    """More candidate pairs inside one tile than the tiled engine's shared-memory pair list holds (32 entries for
    16-cell tiles, 64 for 32-cell tiles): a serpentine chain, every particle with two neighbours at 0.8 cutoff.
    The overflow must fall back to the exact path and stay bit-identical to the oracle."""
    size = box_size(4000)
    pts = []
    for r in range(6):                       # 6 rows, 0.012 apart (rows do not interact), 14 particles each, 0.008 apart
        xs = 0.205 + 0.008 * np.arange(14)
        pts += [(x if r % 2 == 0 else xs[-1] - (x - xs[0]), 0.21 + 0.012 * r) for x in xs]
    parts = np.zeros((len(pts), 6))
    parts[:, :2] = np.array(pts)
    parts[:, 2] = 0.01 * np.cos(np.arange(len(pts)))
    want = parts.copy()
    sim = make_sim(pkg, parts, size, engine)
    for _ in range(4):
        oracle.step(want, size, 1)
        got = sim.step(1).sync().read_particles()
        assert np.array_equal(got, want)
    assert np.abs(want[:, 4:]).max() > 0
    if engine.startswith("tiled"):
        assert sim.info()["reserved_hw_pairs"] > (32 if engine == "tiled16" else 64)
    if engine.startswith("kstep"):
        assert sim.info()["reserved_hw_pairs"] > 32   # more pairs than one warp's list holds
    sim.close()


@pytest.mark.parametrize("engine", ["kstep16", "kstep32", "kstep48", "kstep64"])
def test_kstep_canonical_order_path_in_fused_launches(pkg, oracle, engine):
    """Particles with three or more in-range neighbours INSIDE a fused 3-step launch: the kstep kernel counts a particle's
    in-range pairs in its force word and, from three on, collects the partners from the tile's candidate-pair list and sums them
    in the oracle's (cell-visit rank, x, y) order.  Hexagons at 0.9 cutoff spacing (every rim particle has three neighbours, the
    centre six), stepped in batches of three so that the pairs come from the list built at the first sub-step."""
    size = box_size(4000)
    pts = []
    for cx, cy, rot in [(0.305, 0.412, 0.1), (0.638, 0.642, 0.7), (0.9997, 0.3203, 1.3)]:   # (the second straddles 64-cell tiles)
        pts.append((cx, cy))
        pts += [(cx + 0.009 * np.cos(rot + k * np.pi / 3), cy + 0.009 * np.sin(rot + k * np.pi / 3)) for k in range(6)]
    parts = np.zeros((len(pts), 6))
    parts[:, :2] = np.array(pts)
    parts[:, 2] = 0.05 * np.cos(np.arange(len(pts)))
    parts[:, 3] = 0.05 * np.sin(3 * np.arange(len(pts)))
    want = parts.copy()
    sim = make_sim(pkg, parts, size, engine)
    for _ in range(4):
        oracle.step(want, size, 3)
        got = sim.step(3).sync().read_particles()
        assert np.array_equal(got, want)
    assert oracle.stats(parts, size)["max_neighbours"] >= 6
    info = sim.info()
    assert info["engine"] == pkg.ENGINE_KSTEP and info["engine_switches"] == 0
    sim.close()


def test_kstep_hands_over_when_a_clump_exceeds_its_lists(pkg, oracle):
    """More in-range neighbours per particle than the kstep kernel's force word / canonical-path scratch hold (a 40-particle
    clump half a cutoff wide): the launch reports the overflow, leaves its input intact, and the handle finishes the batch on the
    cellsort engine (which has no capacities) -- still bit-identical to the oracle."""
    rng = np.random.default_rng(3)
    size = box_size(4000)
    n = 40
    parts = np.zeros((n + 50, 6))
    parts[:n, 0] = 0.7 + 0.005 * rng.random(n)
    parts[:n, 1] = 0.7 + 0.005 * rng.random(n)
    parts[n:, 0] = 0.1 + 1.2 * rng.random(50)
    parts[n:, 1] = 0.1 + 1.2 * rng.random(50)
    want = parts.copy()
    sim = pkg.Simulation(parts, len(parts), size, engine=pkg.ENGINE_KSTEP, tile_cells=64)
    assert sim.info()["engine"] == pkg.ENGINE_KSTEP
    oracle.step(want, size, 3)
    got = sim.step(3).sync().read_particles()
    info = sim.info()
    assert info["engine"] == pkg.ENGINE_CELLSORT and info["engine_switches"] == 1
    assert rel_err(got, want) <= 1e-12
    sim.close()


def test_empty_and_dense_inputs(pkg, oracle):
    sim = pkg.Simulation(np.zeros((0, 6)), 0, 0.5)
    sim.step(3).sync()
    assert sim.read_particles().shape == (0, 6)
    sim.close()
    # 600 particles inside one tile: the tiled engine refuses, AUTO falls back to cellsort and matches
    rng = np.random.default_rng(0)
    n, size = 600, box_size(1000)
    parts = np.zeros((n, 6))
    parts[:, 0] = 0.2 + 0.1 * rng.random(n)
    parts[:, 1] = 0.2 + 0.1 * rng.random(n)
    with pytest.raises(pkg.PsimError) as ei:
        pkg.Simulation(parts, n, size, engine=pkg.ENGINE_TILED, tile_cells=16)
    assert ei.value.status == 7
    sim = pkg.Simulation(parts, n, size)
    assert sim.info()["engine"] == pkg.ENGINE_CELLSORT
    want = parts.copy()
    oracle.step(want, size, 3)
    got = sim.step(3).sync().read_particles()
    assert rel_err(got[:, :4], want[:, :4]) <= 1e-12 and np.array_equal(got, want)
    sim.close()


@pytest.mark.parametrize("engine", ["cellsort", "tiled32", "kstep32"])
def test_accelerations_after_a_step_that_did_not_store_them_read_as_zero(pkg, oracle, engine):
    """include/psim.h: ax, ay are those of the LAST step if it stored them, else 0 -- never values that belong to an older
    step or to another particle (the particle order inside the engines changes every step)."""
    n = 5000
    size = box_size(n)
    parts = pkg.init_particles(n, 4)
    oracle.step(parts, size, 60)
    sim = make_sim(pkg, parts, size, engine)
    want = parts.copy()
    oracle.step(want, size, 3)
    got = sim.step(3).sync().read_particles()
    assert np.array_equal(got, want) and np.abs(got[:, 4:]).max() > 0
    oracle.step(want, size, 2)
    got = sim.step(2, pkg.STEP_ACCEL_NONE).sync().read_particles()
    assert np.array_equal(got[:, :4], want[:, :4]) and not got[:, 4:].any()
    oracle.step(want, size, 4)
    got = sim.step(4, pkg.STEP_ACCEL_ALL).sync().read_particles()
    assert np.array_equal(got, want)
    sim.close()


def test_device_pointer_flavour(pkg, oracle):
    """part3/main.cu hands init_simulation a cudaMalloc'ed AoS; read-back into a device array too."""
    import torch

    n = 5000
    size = box_size(n)
    parts = pkg.init_particles(n, 2)
    oracle.step(parts, size, 40)
    dev = torch.from_numpy(parts).cuda()
    sim = pkg.Simulation(dev, n, size)
    sim.step(7).sync()
    out = torch.zeros_like(dev)
    sim.read_particles(out)
    want = parts.copy()
    oracle.step(want, size, 7)
    assert np.array_equal(out.cpu().numpy(), want)
    sim.close()


def test_device_side_generator_and_async_save_path(pkg, oracle):
    """the alternative seeding mode generates the reference's construction on the device (every particle on
    its own lattice site, float velocities in [-1, 1)); a run from it matches the oracle; the asynchronous position read-back
    returns the same frames as the synchronous one while later steps are already enqueued."""
    import torch

    n = 50000
    size = box_size(n)
    dev = torch.empty((n, 6), dtype=torch.float64, device="cuda")
    pkg.generate_particles_device(dev, n, 9, size)
    torch.cuda.synchronize()
    parts = dev.cpu().numpy()
    sx = int(np.ceil(np.sqrt(n)))
    sy = (n + sx - 1) // sx
    col = np.rint(parts[:, 0] * (1 + sx) / size - 1).astype(np.int64)
    row = np.rint(parts[:, 1] * (1 + sy) / size - 1).astype(np.int64)
    site = row * sx + col
    assert site.min() >= 0 and site.max() < n and len(np.unique(site)) == n          # a permutation of the lattice sites
    assert np.abs(np.sort(site) - np.arange(n)).max() == 0 and (site != np.arange(n)).mean() > 0.99   # ... and a shuffled one
    v = parts[:, 2:4]
    assert v.min() >= -1.0 and v.max() < 1.0 and np.array_equal(v, v.astype(np.float32).astype(np.float64))
    assert abs(v.mean()) < 0.02 and abs(v.std() - 1 / np.sqrt(3)) < 0.02 and not parts[:, 4:].any()
    dev2 = torch.empty_like(dev)
    pkg.generate_particles_device(dev2, n, 9, size)
    assert torch.equal(dev, dev2)                                                       # deterministic in (n, seed)
    sim = pkg.Simulation(dev, n, size)                                                  # device-pointer flavour: no upload
    want = parts.copy()
    frames = [torch.empty((n, 2), dtype=torch.float64).pin_memory() for _ in range(2)]
    for k in range(4):
        sim.step(25)
        sim.read_positions_begin(frames[k & 1])
        sim.step(3)                                     # enqueued behind the read: must not disturb the frame
        sim.read_positions_end()
        oracle.step(want, size, 25)
        assert np.array_equal(frames[k & 1].numpy(), want[:, :2]), k
        oracle.step(want, size, 3)
    assert np.array_equal(sim.sync().read_particles()[:, :4], want[:, :4])
    sim.close()


# ---------------------------------------------------------------- slabs over several GPUs
@pytest.mark.parametrize("engine", ["kstep", "tiled"])
@pytest.mark.parametrize("tile", [16, 32])
def test_slabs_match_oracle_on_all_visible_gpus(tile, engine):
    """One process per GPU (torchrun), halo exchange + migration over NCCL, merged state bit-identical to the oracle."""
    import torch

    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least two GPUs")
    world = min(ngpu, 8)
    n = 60000 if tile == 16 else 400000
    cmd = ["python", "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + tile + (1 if engine == "tiled" else 0)), os.path.join(os.path.dirname(__file__), "slab_check.py"), str(n), "120", str(tile), engine]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0 and "SLAB_CHECK ok" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]


# ---------------------------------------------------------------- the reference's multi-process flavour
def _run_mpi_flavour(world, n, steps, port):
    script = os.path.join(os.path.dirname(__file__), "mpi_flavour_check.py")
    if world == 1:
        cmd = ["python", script, str(n), str(steps)]
    else:
        cmd = ["python", "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), script, str(n), str(steps)]
    env = dict(os.environ)
    env["PSIM_RENDEZVOUS"] = f"/tmp/psim_rendezvous_test_{os.getpid()}_{world}"
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert res.returncode == 0 and "MPI_FLAVOUR_CHECK ok" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]


def test_mpi_flavour_api_one_rank():
    """part2/common.h:30-32 through libpsim_mpi_shim.so with num_procs = 1: init / step / gather_for_save vs the oracle."""
    _run_mpi_flavour(1, 20000, 30, 0)


def test_mpi_flavour_api_all_visible_gpus():
    import torch

    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least two GPUs")
    _run_mpi_flavour(min(ngpu, 8), 400000, 30, 29541)


# ---------------------------------------------------------------- drop-in drivers
def _trajectory_md5(cmd, tmp_path, env=None):
    out = tmp_path / "traj.txt"
    e = dict(os.environ)
    e.update(env or {})
    res = subprocess.run(cmd + ["-n", "1000", "-s", "1", "-o", str(out)], capture_output=True, text=True, env=e, timeout=600)
    assert res.returncode == 0, res.stderr
    assert res.stdout.startswith("Simulation Time = ") and res.stdout.rstrip().endswith("seconds for 1000 particles.")
    return hashlib.md5(out.read_bytes()).hexdigest()


def test_own_driver_reproduces_reference_trajectory_file(pkg, tmp_path):
    want = json.load(open(os.path.join(GOLDEN_DIR, "ref_trajectory_n1000_s1.json")))["md5"]
    drv = os.path.join(os.path.dirname(pkg.lib_path()), "psim")
    assert _trajectory_md5([drv], tmp_path) == want
    assert _trajectory_md5([drv], tmp_path, {"PSIM_ENGINE": "cellsort"}) == want


@pytest.mark.parametrize("driver", ["dropin_serial_driver", "dropin_openmp_driver", "dropin_gpu_driver"])
def test_reference_drivers_linked_against_shim(driver, tmp_path):
    """The reference's UNMODIFIED main.cpp / main.cu, linked against libpsim_shim.so instead of the
    reference kernels, must write the same trajectory file as the stock serial binary."""
    path = os.path.join(REF_DIR, driver)
    if not os.path.exists(path):
        pytest.skip("oracle/_ref drop-in drivers not built")
    want = json.load(open(os.path.join(GOLDEN_DIR, "ref_trajectory_n1000_s1.json")))["md5"]
    assert _trajectory_md5([path], tmp_path, {"OMP_NUM_THREADS": "4"}) == want
