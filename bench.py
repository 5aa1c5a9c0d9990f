#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--particles P]
                    [--engine auto|kstep|tiled|cellsort] [--tile 16|32|48|64] [--scaling strong|weak]
                    [--e2e-passes 5] [--no-e2e] [--no-cpu-baseline]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

metric  particle-steps/sec = particles x steps / seconds                     
step    one simulate_one_step over the whole box (bin -> 3x3 force -> move), reference
        part1/serial.cpp:119-131; K steps are enqueued as one psim_step(K) batch
N = 1   20 M particles (configs[3] at one GPU: the configuration the metric is quoted on), `-s 42`
N > 1   the same 20 M particles cut into N slabs of tile rows (strong scaling); --scaling weak runs
        20 M per GPU
value   steps only, state resident in HBM, CUDA events on the launching stream, max over ranks
e2e     psim_create from a PINNED HOST array (H2D inside) + K steps + psim_read_particles back
        into a second pinned host array (D2H inside), wall clock -- the reference driver's own timed
        region (part1/main.cpp:120-144) plus the final read-back; one untimed warm-up pass, then the
        MEDIAN of --e2e-passes passes run back to back (every sample is in the JSON line)
roofline   dominant kernel = the whole step (one fused kernel for the tiled engine); algorithmic
        bytes 80 B per particle-step (read x y vx vy, write x y vx vy ax ay) against the
        measured HBM copy bandwidth in MEASURED_PEAKS.json
cpu_baseline  the UNMODIFIED reference part1/openmp.cpp (oracle/_ref/ref_harness_openmp, built
        by oracle/Makefile) on this box's host cores, a bounded sample (same 20 M particles, few steps)
--impl reference  the reference arm: the same harness, all host threads, same config
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-steps/sec"
ALGO_BYTES = 80.0  # per particle-step, SURVEY.md section 8d
REF_HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness_openmp")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    # untimed steps before the warm-up: the reference's lattice start has no interactions for ~10 steps and thermalises
    # over ~100, so the timed region always measures the steady-state kernel whatever --steps / --warmup say
    ap.add_argument("--prewarm", type=int, default=100)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", type=int, default=20_000_000)
    ap.add_argument("--seed", type=int, default=42)  # the reference's job scripts use -s 42
    ap.add_argument("--engine", default="auto", choices=["auto", "kstep", "tiled", "cellsort"])
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-passes", type=int, default=5, help="timed end-to-end passes (median reported) after one warm-up pass")
    ap.add_argument("--no-bind", action="store_true", help="do not pin the process to the CPUs next to its GPU (A/B knob)")
    ap.add_argument("--cpu-sample-steps", type=int, default=5)
    return ap.parse_args()


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, power, reasons = [], [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        load = [s for s, p in zip(sm, power) if p > 250] or sm
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def bind_near_gpu(torch, local):
    """Pin this process to the CPUs next to its GPU (what `numactl` does in a production launch) so that the pinned host array
    the end-to-end region copies from / to is first touched on the GPU's NUMA node.  Returns the previous affinity mask (restored
    before the CPU baseline runs, which must see every host core), or None when NVML cannot tell."""
    try:
        import pynvml
        before = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        try:   # CUDA_VISIBLE_DEVICES may renumber the devices: go by UUID
            handle = pynvml.nvmlDeviceGetHandleByUUID(f"GPU-{torch.cuda.get_device_properties(local).uuid}".encode())
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(local)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return before
    except Exception:   # no NVML, no permission, unknown topology: run unbound
        return None


def run_reference_harness(n, seed, steps, warmup, threads=None):
    if not os.path.exists(REF_HARNESS):
        raise FileNotFoundError(f"{REF_HARNESS} missing (built by `make -C oracle ref` in the build container)")
    env = dict(os.environ)
    cores = threads or os.cpu_count() or 1
    env.update({"OMP_NUM_THREADS": str(cores), "OMP_PLACES": "cores", "OMP_PROC_BIND": "spread"})  # part1/job-openmp:9-11
    out = subprocess.run([REF_HARNESS, "-n", str(n), "-s", str(seed), "-k", str(steps), "-w", str(warmup)],
                         capture_output=True, text=True, env=env, check=True)
    return json.loads(out.stdout.strip().splitlines()[-1])


def run_oracle_port(n, seed, steps, warmup):
    """Fallback when oracle/_ref is absent (a clean checkout on a box without /root/reference): the oracle's C restatement
    of part1/serial.cpp (oracle/psim_oracle.c, built here by its Makefile), one thread, same particles."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    from psim_testlib import Oracle

    import __graft_entry__ as g

    pkg = g.load_package()
    size = pkg.box_size(n)
    parts = np.zeros((n, 6))
    pkg.init_particles(n, seed, size, out=parts)   # host-side generator of libpsim (no GPU involved)
    orc = Oracle()
    t0 = time.perf_counter()
    orc.step(parts, size, warmup)
    t1 = time.perf_counter()
    orc.step(parts, size, steps)
    t2 = time.perf_counter()
    return {"particle_steps_per_s": n * steps / (t2 - t1), "steps_s": t2 - t1, "init_s": t1 - t0, "threads": 1, "kind": "port"}


def cpu_reference(n, seed, steps, warmup):
    """(result dict, kind, description): the unmodified reference if its binaries travelled with the snapshot, else the port"""
    if os.path.exists(REF_HARNESS):
        r = run_reference_harness(n, seed, steps, warmup)
        return r, "reference", "the unmodified part1/openmp.cpp"
    r = run_oracle_port(n, seed, max(1, min(steps, 3)), 1)
    return r, "port", "oracle/psim_oracle.c (C restatement of part1/serial.cpp; oracle/_ref was not built on this checkout)"


def reference_arm(args):
    """The reference's own CPU implementation (unmodified part1/openmp.cpp) on this box's cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.particles * (args.gpus if args.scaling == "weak" else 1)
    # bounded sample: ~1 s per 20 M-particle step on 16 cores -> cap the timed steps
    k = max(1, min(args.steps, 20))
    w = max(1, min(args.warmup, 3))
    r, kind, what = cpu_reference(n, args.seed, k, w)
    if kind == "port":
        k, w = max(1, min(k, 3)), 1
    value = r["particle_steps_per_s"]
    sample = f"{n} particles, seed {args.seed}, {w} warm-up + {k} timed steps of {what}"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "particle-steps/s", "n_gpus": args.gpus,
        "steps": k, "warmup": w, "ms_per_step": 1e3 * r["steps_s"] / k, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(n, args.seed), "baseline_config": baseline_config(n, args.gpus, args.scaling),
                   "particles": n, "requested_steps": args.steps, "init_simulation_s": r["init_s"]},
        "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": r["threads"], "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_string(n, seed):
    """identical in both arms (the driver compares config.workload)"""
    return f"{n} particles, density 0.0005, cutoff 0.01, seed {seed}, reference generator (part1/main.cpp:31-59)"


def baseline_config(n, world, scaling):
    if scaling == "weak" and world > 1:
        return "configs[4] weak scaling, 20 M particles per GPU"
    return {20_000_000: "configs[3] 20 M particles", 1_000_000: "configs[2] 1 M particles (state fits L2: not an HBM case)",
            100_000: "configs[1] 100 k particles (launch-latency bound)"}.get(n, "-")


def main():
    args = parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__ as g

    pkg = g.load_package()
    pkg.lib()  # fail loudly if the CUDA library is missing: there is no other path

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n = args.particles * (world if args.scaling == "weak" else 1)
    size = pkg.box_size(n)
    affinity_before = None if args.no_bind else bind_near_gpu(torch, local)
    engine = {"auto": pkg.ENGINE_AUTO, "kstep": pkg.ENGINE_KSTEP, "tiled": pkg.ENGINE_TILED, "cellsort": pkg.ENGINE_CELLSORT}[args.engine]

    host = torch.empty((n, 6), dtype=torch.float64, pin_memory=True)
    pkg.init_particles(n, args.seed, size, out=host.numpy())

    def allreduce(value, op, dtype=torch.float64):
        t = torch.tensor([value], dtype=dtype, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=op)
        return t.item()

    def fingerprint(sim):
        """order-independent 64-bit state hash summed over the slabs (mod 2^64) + number of owned particles over all slabs"""
        h, owned = sim.state_hash()
        hs = allreduce(h - (1 << 64) if h >= (1 << 63) else h, dist.ReduceOp.SUM if world > 1 else None, torch.int64)
        return f"{int(hs) & ((1 << 64) - 1):016x}", int(allreduce(owned, dist.ReduceOp.SUM if world > 1 else None, torch.int64))

    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        sim = pkg.Simulation(host, n, size, engine=engine, device=local, stream=stream.cuda_stream,
                             tile_cells=args.tile, rank=rank, nranks=world)
        if world > 1:
            uid = [pkg.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            sim.comm_connect(uid[0])

        def barrier():
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        # ---- device-resident steps -----------------------------------------------------------------
        sim.step(args.prewarm, pkg.STEP_ACCEL_NONE).sync()
        sim.step(args.warmup, pkg.STEP_ACCEL_NONE).sync()
        launches0 = sim.info()["kernel_launches"]
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
            time.sleep(0.3)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(stream)
        sim.step(args.steps, pkg.STEP_DEFAULT)
        ev1.record(stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
        sim.sync()
        clocks = sampler.stop() if rank == 0 else None
        info = sim.info()
        launches = info["kernel_launches"] - launches0
        ms = float(allreduce(ms, dist.ReduceOp.MAX if world > 1 else None))
        value = n * args.steps / (ms * 1e-3)

        # ---- checks on the state that was just timed (assertions, not decoration) ---------------------
        state_hash, owned_total = fingerprint(sim)
        xy = torch.full((n, 2), float("nan"), dtype=torch.float64, device="cuda")
        sim.read_positions(xy)   # a slab writes only the particles it owns
        mine = ~torch.isnan(xy[:, 0])
        inside = bool(((xy[mine] >= 0.0) & (xy[mine] <= size)).all().item())
        inside = bool(allreduce(1.0 if inside else 0.0, dist.ReduceOp.MIN if world > 1 else None) == 1.0)
        del xy, mine
        st = sim.stats()
        pairs = int(allreduce(st["pairs"], dist.ReduceOp.SUM if world > 1 else None, torch.int64))
        ke = float(allreduce(st["kinetic_energy"], dist.ReduceOp.SUM if world > 1 else None))
        dmin = float(allreduce(st["dmin"], dist.ReduceOp.MIN if world > 1 else None))
        vmax = float(allreduce(st["vmax"], dist.ReduceOp.MAX if world > 1 else None))
        recoveries = int(allreduce(info["recoveries"], dist.ReduceOp.SUM if world > 1 else None, torch.int64))
        switches = int(allreduce(info["engine_switches"], dist.ReduceOp.SUM if world > 1 else None, torch.int64))
        steps_done = info["steps_done"]
        if world > 1 and not args.no_e2e:
            # one untimed gather: NCCL sets up its send / receive paths on first use (hundreds of ms, once per process) --
            # library warm-up like the W warm-up steps, not part of the end-to-end region below
            sim.gather(host if rank == 0 else None, root=0)
        sim.close()
        assert owned_total == n, f"the slabs own {owned_total} particles, expected {n}"
        assert inside, "a particle left the box"
        assert steps_done == args.prewarm + args.warmup + args.steps, steps_done
        assert 0.5 < dmin <= 1.0 and vmax < 20.0, (dmin, vmax)

        # ---- end to end through the C ABI with host buffers -----------------------------------------
        # One pass = psim_create from the pinned host array (H2D inside) + K steps + read-back into a second pinned host array
        # (D2H inside), wall clock.  The passes run back to back: one untimed warm-up pass, then --e2e-passes timed ones; the
        # MEDIAN pass is reported and every sample is listed (a shared host's PCIe / memory traffic moves a single create
        # between 24 and 200 ms on this pool: profiles/tools/e2e_gap_test.py).
        e2e = None
        if not args.no_e2e:
            pkg.init_particles(n, args.seed, size, out=host.numpy())
            host_out = torch.empty((n, 6), dtype=torch.float64, pin_memory=True) if rank == 0 else None
            samples, phases, e2e_hash, e2e_owned = [], None, None, None
            for it in range(1 + max(1, args.e2e_passes)):
                barrier()
                t0 = time.perf_counter()
                sim2 = pkg.Simulation(host, n, size, engine=engine, device=local, stream=stream.cuda_stream,
                                      tile_cells=args.tile, rank=rank, nranks=world)
                t1 = time.perf_counter()
                if world > 1:
                    sim2.comm_connect(uid[0])   # same id: the process-level communicator is reused (like MPI_COMM_WORLD)
                t2 = time.perf_counter()
                sim2.step(args.steps, pkg.STEP_DEFAULT)
                sim2.sync()
                t3 = time.perf_counter()
                if world > 1:
                    sim2.gather(host_out, root=0)   # rank 0 ends up with ALL particles in original order in its pinned
                else:                               # host array (the reference's gather_for_save)
                    sim2.read_particles(host_out)
                barrier()
                t4 = time.perf_counter()
                dt = float(allreduce(t4 - t0, dist.ReduceOp.MAX if world > 1 else None))
                if it == 0:   # warm-up pass: also the place to check what came back
                    e2e_hash, e2e_owned = fingerprint(sim2)
                    if rank == 0:
                        back = host_out.numpy()
                        assert np.isfinite(back).all() and (back[:, :2] >= 0.0).all() and (back[:, :2] <= size).all(), "read-back outside the box"
                else:
                    samples.append(dt)
                    if phases is None or dt <= min(samples):
                        phases = {"create_h2d": t1 - t0, "connect": t2 - t1, "steps": t3 - t2, "read_back_d2h": t4 - t3}
                sim2.close()
            assert e2e_owned == n
            dt = statistics.median(samples)
            e2e = {"value": n * args.steps / dt, "unit": "particle-steps/s", "seconds": dt,
                   "h2d_bytes_per_step": 48.0 * n / args.steps, "d2h_bytes_per_step": 48.0 * n / args.steps,
                   "passes": {"warmup": 1, "timed": len(samples), "reported": "median", "seconds_all": samples,
                              "phases_s_rank0_fastest_pass": phases},
                   "state_hash_after_init_plus_steps": e2e_hash,
                   "what": ("psim_create(pinned host AoS, H2D inside) + psim_step(K) + psim_read_particles(-> pinned host AoS, D2H inside), "
                            "wall clock" if world == 1 else
                            "every rank: psim_create(the same pinned host AoS) + psim_comm_connect (each rank uploads 1/N of the array, "
                            "the records reach their slabs over NVLink) + psim_step(K) + psim_gather(all particles, original order -> rank "
                            "0's pinned host AoS, D2H inside), wall clock, max over ranks; the NCCL communicator and its send / receive "
                            "paths were warmed by the timed run and the warm-up pass (like MPI_Init before the reference's timer)")}
            del host_out

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    per_gpu_particles = n / world
    kstep = info["engine"] == pkg.ENGINE_KSTEP
    k = max(1, info["steps_per_launch"]) if kstep else 1
    step_launches = max(1, round(launches * k / args.steps))  # kernels per batch of k steps on this rank
    achieved = ALGO_BYTES * per_gpu_particles / (ms * 1e-3 / args.steps) / 1e9
    engine_name = {pkg.ENGINE_TILED: "tiled", pkg.ENGINE_KSTEP: "kstep"}.get(info["engine"], "cellsort")
    if kstep:
        ts, h = info["tile_cells"], info["halo_cells"]
        region = (ts + 2 * h) ** 2 / float(ts * ts)
        # per launch of k steps: read x y vx vy of the region (tile + halo), ids of the tile; write x y vx vy id of the tile
        moved = (32.0 * region + 4.0 + 36.0) / k
        kernel = f"kstep_kernel<{ts},{h}> ({k} steps fused per launch)"
    else:
        moved = 72.0 if engine_name == "tiled" else None
        kernel = "tile_step_kernel" if engine_name == "tiled" else "hist+scan+scatter+force_move (4 kernels)"
    line = {
        "metric": METRIC, "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(n, args.seed), "baseline_config": baseline_config(n, world, args.scaling),
                   "particles": n, "engine": engine_name, "tile_cells": info["tile_cells"], "halo_cells": info["halo_cells"],
                   "steps_per_launch": info["steps_per_launch"], "slabs": world,
                   "untimed_steps_before": args.prewarm + args.warmup,
                   "l2": "state (>= 640 MB per GPU at 20 M) larger than the 126 MB L2; no flush needed" if per_gpu_particles >= 4e6
                         else "state fits L2: the HBM roofline fraction is not meaningful at this size",
                   "accel_store": "ax, ay materialised for the last step of the batch (the reference drivers never read them)",
                   "device_bytes": info["device_bytes"],
                   "host_affinity": "bound to the GPU's NUMA node (NVML) while the pinned host array is allocated and copied"
                                    if affinity_before is not None else "unbound"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None,
                     "algorithmic_bytes_per_particle_step": ALGO_BYTES,
                     "algorithmic_bytes_per_launch": ALGO_BYTES * per_gpu_particles * k / step_launches,
                     "bytes_moved_per_particle_step_model": moved,
                     "frac_of_peak_on_bytes_moved": (moved * per_gpu_particles / (ms * 1e-3 / args.steps) / 1e9 / peak) if moved else None,
                     "kernel": kernel, "launches_per_batch": step_launches, "steps_per_launch": k, "peak_source": peak_src,
                     "note": "achieved = 80 algorithmic bytes x particle-steps/s (SURVEY 8d). The kstep kernel keeps a tile K steps in "
                             "shared memory, so it MOVES fewer bytes than that (model above; ncu dram bytes under profiles/) and is bound "
                             "by shared-memory / issue throughput, not HBM. traffic is null here: it is only quoted from an ncu capture."},
        "clocks": clocks,
        "check": {"state_hash": state_hash, "owned_particles": owned_total, "all_inside_box": inside, "steps_done": steps_done,
                  "pairs": pairs, "dmin": dmin, "vmax": vmax, "kinetic_energy": ke, "speed_bound_replays": recoveries,
                  "engine_switches": switches,
                  "note": "state_hash = sum over all slabs of a 64-bit mix of (id, x, y, vx, vy): identical for 1/2/4/8 GPUs iff the "
                          "states are bit-identical; asserted: every particle owned exactly once and inside the box; pairs / dmin / "
                          "kinetic energy are reduced over the slabs (a slab also sees its ghost rows as neighbours)"},
    }
    if e2e:
        line["e2e"] = e2e
    if affinity_before is not None:
        os.sched_setaffinity(0, affinity_before)   # the CPU baseline below uses every host core
    if world == 1 and not args.no_cpu_baseline:
        try:
            r, kind, what = cpu_reference(n, args.seed, args.cpu_sample_steps, 1)
            line["cpu_baseline"] = {
                "value": r["particle_steps_per_s"], "unit": "particle-steps/s", "cores": r["threads"], "kind": kind,
                "sample": f"{n} particles, seed {args.seed}, 1 warm-up + {args.cpu_sample_steps if kind == 'reference' else min(args.cpu_sample_steps, 3)} "
                          f"timed steps of {what} (init / warm-up {r['init_s']:.1f} s not counted)"}
        except Exception as exc:  # the bench line must still appear
            line["cpu_baseline"] = {"value": None, "unit": "particle-steps/s", "cores": os.cpu_count(), "kind": "reference",
                                    "sample": f"failed: {exc}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
