"""Regenerates the committed fixtures in tests/golden/ FROM THE UNMODIFIED REFERENCE.

Run in the build container only (needs `make -C oracle ref`, i.e. /root/reference):

    python tests/golden/gen_golden.py

The reference ships no golden vectors for its hot path, so the pins are
manufactured from the reference's own code: part1/main.cpp's generator (init_particles) and
part1/serial.cpp's init_simulation / simulate_one_step, driven in memory through
oracle/_ref/libref_{driver,serial}.so (full double precision -- the driver's text output keeps only
six digits), plus the md5 of the stock `serial -n 1000 -s 1 -o f` trajectory file
(BASELINE.json configs[0]).

Files written
  ref_n1000_s1.npz   states (N,6) after 0, 1, 50, 51, 100, 300, 1000 reference steps; per-cell counts
                     and per-particle cell ids read from the reference's own `Bins` at steps 0/100/1000
  ref_n3000_s7.npz   same at steps 0, 100, 101, 200 (a second seed and a non-square particle count)
  ref_trajectory_n1000_s1.json   md5 + head/tail lines of the stock trajectory file
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from psim_testlib import REF_DIR, RefKernel, box_size, ref_init_particles  # noqa: E402


def cell_ids_from_lists(start, member, n):
    ids = np.empty(n, dtype=np.int32)
    for c in np.nonzero(np.diff(start))[0]:
        ids[member[start[c]:start[c + 1]]] = c
    return ids


def states(n, seed, marks, cell_marks):
    size = box_size(n)
    parts = ref_init_particles(n, seed)
    out = {"n": np.int64(n), "seed": np.int64(seed), "size": np.float64(size)}
    ref = RefKernel("serial").init(parts, size)
    done = 0
    for m in marks:
        ref.step(m - done)
        done = m
        out[f"step{m}"] = parts.copy()
        if m in cell_marks:
            start, member = ref.cell_lists()
            out[f"cellcount{m}"] = ref.cell_counts()
            out[f"cellid{m}"] = cell_ids_from_lists(start, member, n)
    out["bincnt"] = np.int64(ref.bin_count())
    return out


def main():
    np.savez_compressed(os.path.join(HERE, "ref_n1000_s1.npz"),
                        **states(1000, 1, [0, 1, 50, 51, 100, 300, 1000], {0, 100, 1000}))
    np.savez_compressed(os.path.join(HERE, "ref_n3000_s7.npz"),
                        **states(3000, 7, [0, 100, 101, 200], {0, 100, 200}))
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "traj.txt")
        subprocess.check_call([os.path.join(REF_DIR, "ref_serial"), "-n", "1000", "-s", "1", "-o", f],
                              stdout=subprocess.DEVNULL)
        data = open(f, "rb").read()
    lines = data.decode().split("\n")
    meta = {"command": "serial -n 1000 -s 1 -o f  (reference part1/main.cpp + part1/serial.cpp, g++ -O3 -std=c++11)",
            "md5": hashlib.md5(data).hexdigest(), "bytes": len(data), "lines": len(lines) - 1,
            "head": lines[:4], "tail": lines[-4:]}
    json.dump(meta, open(os.path.join(HERE, "ref_trajectory_n1000_s1.json"), "w"), indent=1)
    print(meta["md5"], meta["lines"])


if __name__ == "__main__":
    main()
