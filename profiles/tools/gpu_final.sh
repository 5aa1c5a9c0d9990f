#!/bin/bash
# Round-end evidence on one box: full GPU test-suite, smoke, bench (driver flags; defaults with the CPU baseline), reference arm,
# then the profiler passes (launch list of bench.py, one `ncu --set full` capture of a steady-state kstep launch).
TAG=${1:-r2h}
B=parallel-particle-simulation_b200/csrc/build
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_driver_flags.json 2> gpurun_out/${TAG}_bench_driver_flags.err; echo "bench(driver flags) rc=$?"
python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err; echo "bench(default) rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_reference_arm.err; echo "reference arm rc=$?"
ncu --set full --clock-control none --import-source on -k regex:kstep_kernel --launch-skip 60 --launch-count 1 -f -o gpurun_out/${TAG}_kstep $B/psim -n 20000000 -s 42 > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python - <<PY
import json
for f in ("bench_driver_flags", "bench_default", "bench_reference_arm"):
    try:
        d = json.load(open("gpurun_out/${TAG}_%s.json" % f))
        print(f, "%.3f G" % (d["value"] / 1e9), "ms/step %.4f" % d["ms_per_step"], "e2e %.3f G" % (d["e2e"]["value"] / 1e9), d.get("check", {}).get("state_hash"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "FAILED", e)
PY
