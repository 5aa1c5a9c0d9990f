"""Host-side logic that needs no GPU: the slab decomposition across ranks (checked by two real processes over gloo),
the reference arm's rank gating under a multi-process launch, and the pin of the kernels' division-free cell index."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, %(root)r)
import torch, torch.distributed as dist
import __graft_entry__ as g
pkg = g.load_package()
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
out = {}
for n in (1000, 100000, 20000000, 160000000):
    nb = pkg.bin_count(pkg.box_size(n))
    for tile in (0, 16, 32, 64):
        try:
            b, e = pkg.slab_rows(nb, rank, world, tile)
        except pkg.PsimError as ex:
            b, e = -1, -1                      # more slabs than tile rows
        t = torch.tensor([b, e], dtype=torch.int64)
        got = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(got, t)                # every rank sees the whole decomposition
        out[f"{n}/{tile}"] = [[int(x[0]), int(x[1])] for x in got] + [nb]
# particle ownership: every particle of a generated state falls into exactly one slab
n = 20000
parts = pkg.init_particles(n, 3)
nb = pkg.bin_count(pkg.box_size(n))
b, e = pkg.slab_rows(nb, rank, world, 16)
rows = np.minimum(np.floor(parts[:, 0] / 0.01).astype(np.int64), nb - 1)
mine = torch.from_numpy(((rows >= b) & (rows < e)).astype(np.int64))
dist.all_reduce(mine)
out["owned_once"] = bool((mine == 1).all())
if rank == 0:
    print("RESULT " + json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
'''


def test_slab_decomposition_two_ranks_gloo(tmp_path):
    """world_size 2 over gloo: both ranks compute their slab through the C ABI; together the slabs tile [0, bin_count)
    without gap or overlap, on tile boundaries, for every BASELINE configuration and tile size."""
    script = tmp_path / "worker.py"
    script.write_text("import numpy as np\n" + WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29641")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29641", str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("RESULT ")][-1]
    out = json.loads(line[len("RESULT "):])
    assert out.pop("owned_once") is True
    for key, val in out.items():
        nb, slabs = val[-1], val[:-1]
        if slabs[0][0] < 0:
            continue
        tile = int(key.split("/")[1])
        assert slabs[0][0] == 0 and slabs[-1][1] == nb, key
        for (b0, e0), (b1, e1) in zip(slabs, slabs[1:]):
            assert e0 == b1 and b0 < e0 and b1 < e1, key
            if tile:
                assert e0 % tile == 0, key


def test_slab_rows_rejects_bad_arguments(pkg):
    with pytest.raises(pkg.PsimError):
        pkg.slab_rows(100, 3, 2)
    with pytest.raises(pkg.PsimError):
        pkg.slab_rows(100, 0, 2, 40)     # not a tile size of this build (16, 32, 48, 64)
    with pytest.raises(pkg.PsimError):
        pkg.slab_rows(71, 0, 9, 16)      # 5 tile rows cannot feed 9 slabs
    assert pkg.slab_rows(71, 0, 1, 16) == (0, 71)
    assert pkg.slab_rows(100, 0, 2, 48) == (0, 96)   # 3 tile rows of 48 cells: two for rank 0, the (partial) third for rank 1
    assert pkg.slab_rows(100, 1, 2, 48) == (96, 100)


def test_reference_arm_only_rank0_works(tmp_path):
    """bench.py --impl reference under a multi-process launch: ranks other than 0 exit 0 without output or work."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, env=env, timeout=120)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_division_free_cell_index_is_pinned():
    """oracle/check_div: the kernels' two-FMA quotient equals the IEEE division on every double near every cell edge
    (16 M values) and on random positions (sample reduced here; the full 2e9 run is recorded in DESIGN.md)."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "check_div"])
    res = subprocess.run([os.path.join(ROOT, "oracle", "check_div"), "5000000"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "0 mismatches" in res.stdout, res.stdout


def test_oracle_cells_agree_with_fma_quotient(oracle):
    """The same quotient in numpy-free Python on the oracle's own generated state: floor(q) == the oracle's cell ids."""
    import math
    from psim_testlib import box_size
    import __graft_entry__ as g
    pkg = g.load_package()
    n = 20000
    parts = pkg.init_particles(n, 9)
    size = box_size(n)
    oracle.step(parts, size, 25)
    nb = pkg.bin_count(size)
    ids = oracle.cell_ids(parts, size)
    def q(x):
        q0 = x * 100.0
        r = math.fma(-q0, 0.01, x) if hasattr(math, "fma") else None
        return None if r is None else math.fma(r, 100.0, q0)
    if q(0.5) is None:
        pytest.skip("math.fma needs Python 3.13")
    mine = np.array([min(int(math.floor(q(x))), nb - 1) * nb + min(int(math.floor(q(y))), nb - 1) for x, y in parts[:, :2]])
    assert np.array_equal(mine, ids)
