#!/bin/bash
set -u
mkdir -p gpurun_out/r2r
O=gpurun_out/r2r
( time timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err ) 2> $O/time.txt; echo "rc=$?"; cat $O/time.txt | tail -3
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2r/bench_default.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["cpu_baseline"]["value"])
for k,v in d["other_workloads"].items(): print(k, {a:v.get(a) for a in ("value","unit","ms_per_step","wall_s","error")}, v.get("roofline",{}).get("frac"))
PY
