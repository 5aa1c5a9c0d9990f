for i in 1 2; do
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab_bind_$i.json 2>gpurun_out/ab_bind_$i.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-bind > gpurun_out/ab_nobind_$i.json 2>gpurun_out/ab_nobind_$i.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/ab_*.json")):
    d=json.load(open(f)); print(f, "%.2f G"%(d["value"]/1e9), "e2e %.2f G"%(d["e2e"]["value"]/1e9), {k:round(v*1e3,1) for k,v in d["e2e"]["phases_s_rank0"].items()}, d["config"]["host_affinity"][:10])
PY
nvidia-smi topo -m | head -8; numactl -H 2>/dev/null | head -5; lscpu | grep -i numa
