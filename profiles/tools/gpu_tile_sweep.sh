#!/bin/bash
# A/B of kstep tile sizes (variant builds under csrc/build_t*/, see PSIM_KSTEP_TSX in psim_kstep.cu): one short bench line each.
# Every line carries check.state_hash: identical hashes = bit-identical 20 M-particle states after 720 steps.
C=parallel-particle-simulation_b200/csrc
mkdir -p gpurun_out
run() { # tag lib tile
  PSIM_LIB=$PWD/$C/$2/libpsim.so python bench.py --tile $3 --steps ${STEPS:-600} --warmup 20 --no-cpu-baseline --no-e2e $EXTRA_ARGS > gpurun_out/sweep_$1.json 2> gpurun_out/sweep_$1.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/sweep_$1.json"))
    print("$1", "tile", d["config"]["tile_cells"], "ms/step %.4f" % d["ms_per_step"], "G p-s/s %.2f" % (d["value"] / 1e9), d["check"]["state_hash"], d["check"]["speed_bound_replays"], d["check"]["engine_switches"])
except Exception as e:
    print("$1 FAILED", e, open("gpurun_out/sweep_$1.err").read()[-400:])
PY
}
for spec in "$@"; do IFS=: read tag lib tile <<< "$spec"; run $tag $lib $tile; done
