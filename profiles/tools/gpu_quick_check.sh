#!/bin/bash
# Quick look at a changed kernel: smoke (parity against the oracle on all engines), the kstep parity tests, one short bench line.
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5; echo "smoke rc=$?"
timeout 300 python -m pytest tests -m gpu -x -q -k "kstep or trajectory or short_trajectory or one_step or edge or overflow or clump or 1m_particles" 2>&1 | tail -5
PSIM_TRACE=1 timeout 200 python bench.py --steps 600 --warmup 20 --no-cpu-baseline --no-e2e > gpurun_out/quick.json 2> gpurun_out/quick.err; grep "tiles, halo" gpurun_out/quick.err | head -1
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/quick.json"))
    print("bench: %.2f G p-s/s, %.4f ms/step, hash %s (600+120 steps: b7c5daae49047189), replays %d switches %d" % (d["value"] / 1e9, d["ms_per_step"], d["check"]["state_hash"], d["check"]["speed_bound_replays"], d["check"]["engine_switches"]))
except Exception as e:
    print("bench FAILED", e, open("gpurun_out/quick.err").read()[-800:])
PY
