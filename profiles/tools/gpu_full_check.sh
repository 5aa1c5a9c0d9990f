#!/bin/bash
# The round-end sequence on one box: GPU test-suite, smoke, the bench line with the driver's flags and with the defaults.
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_driver_flags.json 2> gpurun_out/bench_driver_flags.err; echo "bench(driver flags) rc=$?"
python bench.py --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench(default) rc=$?"
python - <<'PY'
import json
for f in ("bench_driver_flags", "bench_default"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "%.2f G" % (d["value"] / 1e9), "ms/step %.4f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "e2e %.2f G" % (d["e2e"]["value"] / 1e9),
              [round(x * 1e3, 1) for x in d["e2e"]["passes"]["seconds_all"]], {k: round(v * 1e3, 1) for k, v in d["e2e"]["passes"]["phases_s_rank0_fastest_pass"].items()},
              d["check"]["state_hash"], d.get("cpu_baseline", {}).get("value"))
    except Exception as e:
        print(f, "FAILED", e, open(f"gpurun_out/{f}.err").read()[-600:])
PY
