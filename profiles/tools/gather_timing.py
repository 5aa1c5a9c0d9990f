#!/usr/bin/env python
"""How long do psim_gather / slab read-back take, cold and warm?  (torchrun, one process per GPU)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import __graft_entry__ as g
pkg = g.load_package()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
size = pkg.box_size(n)
host = torch.empty((n, 6), dtype=torch.float64, pin_memory=True)
pkg.init_particles(n, 42, size, out=host.numpy())
uid = [pkg.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
for rep in range(2):
    t0 = time.perf_counter()
    sim = pkg.Simulation(host, n, size, device=local, rank=rank, nranks=world)
    t1 = time.perf_counter()
    sim.comm_connect(uid[0])
    t2 = time.perf_counter()
    sim.step(20).sync()
    t3 = time.perf_counter()
    times = []
    for k in range(3):
        torch.cuda.synchronize(); dist.barrier()
        a = time.perf_counter()
        sim.gather(host if rank == 0 else None, 0)
        times.append(time.perf_counter() - a)
    torch.cuda.synchronize(); dist.barrier()
    a = time.perf_counter()
    sim.read_particles(host)
    own = time.perf_counter() - a
    sim.close()
    if rank == 0:
        print(f"rep {rep}: create {t1-t0:.4f} connect {t2-t1:.4f} steps {t3-t2:.4f} gather x3 {[round(t,4) for t in times]} own-slab read {own:.4f}", flush=True)
dist.barrier()
dist.destroy_process_group()
