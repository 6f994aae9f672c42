mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2l.json 2> gpurun_out/bench_r2l.err
tail -3 gpurun_out/bench_r2l.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2l.json"))
print({k:d[k] for k in ("value","ms_per_step","e2e","clocks","energy_drift")})
r=d["roofline"]; print({k:r[k] for k in r if k not in ("canonical","source","peak_source")})
print(d["cpu_baseline"])
for k,v in d.get("extra",{}).items(): print(k, {kk:vv for kk,vv in v.items() if kk not in ("config","note")})
PY
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
