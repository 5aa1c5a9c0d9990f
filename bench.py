#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--particles P]
                    [--engine auto|tiled|cellsort] [--tile 16|32|64] [--scaling strong|weak]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

metric  particle-steps/sec = particles x steps / seconds                     
step    one simulate_one_step over the whole box (bin -> 3x3 force -> move), reference
        part1/serial.cpp:119-131; K steps are enqueued as one psim_step(K) batch
N = 1   20 M particles (configs[3] at one GPU: the configuration the metric is quoted on), `-s 42`
N > 1   the same 20 M particles cut into N slabs of tile rows (strong scaling); --scaling weak runs
        20 M per GPU
value   steps only, state resident in HBM, CUDA events on the launching stream, max over ranks
e2e     psim_create from a PINNED HOST array (H2D inside) + K steps + psim_read_particles back
        into the host array (D2H inside), wall clock -- the reference driver's own timed region
        (part1/main.cpp:120-144) plus the final read-back
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
    ap.add_argument("--cpu-sample-steps", type=int, default=5)
    return ap.parse_args()


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any"""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


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


def run_reference_harness(n, seed, steps, warmup, threads=None):
    if not os.path.exists(REF_HARNESS):
        raise FileNotFoundError(f"{REF_HARNESS} missing (built by `make -C oracle ref` in the build container)")
    env = dict(os.environ)
    cores = threads or os.cpu_count() or 1
    env.update({"OMP_NUM_THREADS": str(cores), "OMP_PLACES": "cores", "OMP_PROC_BIND": "spread"})  # part1/job-openmp:9-11
    out = subprocess.run([REF_HARNESS, "-n", str(n), "-s", str(seed), "-k", str(steps), "-w", str(warmup)],
                         capture_output=True, text=True, env=env, check=True)
    return json.loads(out.stdout.strip().splitlines()[-1])


def reference_arm(args):
    """The reference's own CPU implementation (unmodified part1/openmp.cpp) on this box's cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.particles * (args.gpus if args.scaling == "weak" else 1)
    # bounded sample: ~1 s per 20 M-particle step on 16 cores -> cap the timed steps
    k = max(1, min(args.steps, 20))
    w = max(1, min(args.warmup, 3))
    r = run_reference_harness(n, args.seed, k, w)
    value = r["particle_steps_per_s"]
    sample = f"{n} particles, seed {args.seed}, {w} warm-up + {k} timed steps of the unmodified part1/openmp.cpp"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "particle-steps/s", "n_gpus": args.gpus,
        "steps": k, "warmup": w, "ms_per_step": 1e3 * r["steps_s"] / k, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{n} particles, density 0.0005, cutoff 0.01, seed {args.seed}", "particles": n,
                   "requested_steps": args.steps, "init_simulation_s": r["init_s"]},
        "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": r["threads"], "kind": "reference",
                         "sample": sample},
        "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


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
    engine = {"auto": pkg.ENGINE_AUTO, "kstep": pkg.ENGINE_KSTEP, "tiled": pkg.ENGINE_TILED, "cellsort": pkg.ENGINE_CELLSORT}[args.engine]

    host = torch.empty((n, 6), dtype=torch.float64, pin_memory=True)
    pkg.init_particles(n, args.seed, size, out=host.numpy())

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
        launches = sim.info()["kernel_launches"] - launches0
        info = sim.info()
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        value = n * args.steps / (ms * 1e-3)

        # sanity on the state that was just timed: every particle still inside the box
        st = sim.stats()
        sim.close()

        # ---- end to end through the C ABI with host buffers -----------------------------------------
        e2e = None
        if not args.no_e2e:
            pkg.init_particles(n, args.seed, size, out=host.numpy())
            barrier()
            t0 = time.perf_counter()
            sim2 = pkg.Simulation(host, n, size, engine=engine, device=local, stream=stream.cuda_stream,
                                  tile_cells=args.tile, rank=rank, nranks=world)
            if world > 1:
                sim2.comm_connect(uid[0])   # same id: the process-level communicator is reused (like MPI_COMM_WORLD)
            sim2.step(args.steps, pkg.STEP_DEFAULT)
            sim2.read_particles(host)
            barrier()
            dt = time.perf_counter() - t0
            te = torch.tensor([dt], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            dt = float(te.item())
            sim2.close()
            e2e = {"value": n * args.steps / dt, "unit": "particle-steps/s", "seconds": dt,
                   "h2d_bytes_per_step": 48.0 * n / args.steps, "d2h_bytes_per_step": 48.0 * n / args.steps,
                   "what": "psim_create(pinned host AoS) + psim_step(K) + psim_read_particles(host AoS), wall clock"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    per_gpu_particles = n / world
    launches_per_step = max(1, round(launches / args.steps))
    achieved = ALGO_BYTES * per_gpu_particles / (ms * 1e-3 / args.steps) / 1e9
    traffic = recorded_traffic()
    line = {
        "metric": METRIC, "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{n} particles, density 0.0005, cutoff 0.01, seed {args.seed} "
                               f"(BASELINE configs[{4 if args.scaling == 'weak' and world > 1 else 3 if n == 20_000_000 else 2 if n == 1_000_000 else 1 if n == 100_000 else '-'}])",
                   "particles": n, "engine": {pkg.ENGINE_TILED: "tiled", pkg.ENGINE_KSTEP: "kstep"}.get(info["engine"], "cellsort"), "steps_per_launch": info["steps_per_launch"], "halo_cells": info["halo_cells"], "recoveries": info["recoveries"],
                   "tile_cells": info["tile_cells"], "slabs": world, "l2": "state (>= 640 MB per GPU at 20 M) larger than L2; no flush needed",
                   "accel_store": "last step of the batch", "device_bytes": info["device_bytes"]},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": (traffic or {}).get("dram_bytes_per_launch"),
                     "algorithmic_bytes_per_launch": ALGO_BYTES * per_gpu_particles / launches_per_step,
                     "kernel": "tile_step_kernel" if info["engine"] == pkg.ENGINE_TILED else "hist+scan+scatter+force_move (4 kernels)",
                     "launches_per_step": launches_per_step, "peak_source": peak_src},
        "clocks": clocks,
        "check": {"pairs": st["pairs"], "dmin": st["dmin"], "davg": st["davg"], "kinetic_energy": st["kinetic_energy"]},
    }
    if e2e:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu_baseline:
        try:
            r = run_reference_harness(n, args.seed, args.cpu_sample_steps, 1)
            line["cpu_baseline"] = {
                "value": r["particle_steps_per_s"], "unit": "particle-steps/s", "cores": r["threads"], "kind": "reference",
                "sample": f"{n} particles, seed {args.seed}, 1 warm-up + {args.cpu_sample_steps} timed steps of the unmodified "
                          f"part1/openmp.cpp (init_simulation {r['init_s']:.1f} s not counted)"}
        except Exception as exc:  # the bench line must still appear
            line["cpu_baseline"] = {"value": None, "unit": "particle-steps/s", "cores": os.cpu_count(), "kind": "reference",
                                    "sample": f"failed: {exc}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
