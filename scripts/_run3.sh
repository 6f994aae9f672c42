mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err
tail -3 gpurun_out/bench_r2i.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2i.json"))
print({k:d[k] for k in ("value","ms_per_step","e2e","roofline","clocks")})
for k,v in d.get("extra",{}).items(): print(k, {kk:vv for kk,vv in v.items() if kk!="config"})
PY
