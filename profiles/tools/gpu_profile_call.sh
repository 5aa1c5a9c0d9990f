#!/bin/bash
# One GPU-box call: smoke, a bench line, the ncu launch list of bench.py and one `ncu --set full` capture of a steady-state
# kstep launch (through the `psim` driver: no Python start-up under the profiler).  Outputs under gpurun_out/.
TAG=${1:-r2f}
B=parallel-particle-simulation_b200/csrc/build
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
( time python bench.py --steps 600 --warmup 20 --no-cpu-baseline ) > gpurun_out/${TAG}_bench_600.json 2> gpurun_out/${TAG}_bench_600.err; echo "bench rc=$?"
( time ncu --set full --clock-control none --import-source on -k regex:kstep_kernel --launch-skip 60 --launch-count 1 -f -o gpurun_out/${TAG}_kstep \
    $B/psim -n 20000000 -s 42 ) > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
( time ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline ) > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
tail -c 600 gpurun_out/${TAG}_bench_600.json
