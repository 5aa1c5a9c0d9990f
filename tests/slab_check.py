#!/usr/bin/env python
"""Multi-GPU slab parity check, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/slab_check.py [n] [steps] [tile] [kstep|tiled]

Every rank builds its slab of the same warmed state, all ranks step together (halo exchange + migration over
NCCL), every rank writes the particles it owns into its copy of the array, the copies are merged on rank 0 and
compared BIT FOR BIT with the oracle's trajectory.  Prints one line `SLAB_CHECK ok ...` on success."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist

    import __graft_entry__ as g
    from psim_testlib import Oracle

    pkg = g.load_package()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 120
    tile = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    engine_name = sys.argv[4] if len(sys.argv) > 4 else "kstep"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    size = pkg.box_size(n)
    parts = pkg.init_particles(n, 7)
    orc = Oracle()
    orc.step(parts, size, 40)   # warmed state, identical on every rank
    engine = pkg.ENGINE_KSTEP if engine_name == "kstep" else pkg.ENGINE_TILED
    sim = pkg.Simulation(parts.copy(), n, size, engine=engine, tile_cells=tile, device=local, rank=rank, nranks=world)
    uid = [pkg.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    sim.comm_connect(uid[0])
    mine = np.full((n, 6), np.nan)
    done = 0
    for chunk in (1, steps // 2, steps - 1 - steps // 2):
        sim.step(chunk).sync()
        done += chunk
    sim.read_particles(mine)
    owned = ~np.isnan(mine[:, 0])
    t = torch.from_numpy(np.nan_to_num(mine, nan=0.0)).cuda()
    c = torch.from_numpy(owned.astype(np.int32)).cuda()
    dist.reduce(t, 0)   # every particle is owned by exactly one rank: the sum is the merge
    dist.reduce(c, 0)
    info = sim.info()
    st = sim.stats()
    sums = torch.tensor([st["pairs"], st["touched"]], dtype=torch.int64, device="cuda")
    ke = torch.tensor([st["kinetic_energy"]], dtype=torch.float64, device="cuda")
    dmin = torch.tensor([st["dmin"]], dtype=torch.float64, device="cuda")
    dist.reduce(sums, 0)
    dist.reduce(ke, 0)
    dist.reduce(dmin, 0, op=dist.ReduceOp.MIN)
    sim.close()
    ok = True
    if rank == 0:
        want = parts.copy()
        orc.step(want, size, done)
        got = t.cpu().numpy()
        counts = c.cpu().numpy()
        once = bool((counts == 1).all())
        same = bool(np.array_equal(got, want))
        # per-slab statistics count the pairs that straddle a slab border on both sides (ghost rows are neighbours): the sums
        # over the slabs must equal the oracle's numbers for the whole box
        ost = orc.stats(want, size)
        stats_ok = (engine_name != "kstep") or (int(sums[0]) == ost["pairs"] and int(sums[1]) == ost["touched"]
                                                 and abs(float(dmin[0]) - ost["dmin"]) <= 1e-15
                                                 and abs(float(ke[0]) - ost["ke"]) <= 1e-9 * ost["ke"])
        ok = once and same and stats_ok
        print(f"SLAB_CHECK {'ok' if ok else 'FAIL'} engine={engine_name} ranks={world} n={n} steps={done} tile={tile} owned_once={once} "
              f"bit_identical={same} stats_equal_oracle={stats_ok} max_abs_diff={np.abs(got - want).max():.3e} rows={info['row_begin']}..{info['row_end']}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
