#!/bin/bash
# round 2, call AL: final check of the committed state - whole GPU suite, smoke, default bench
set -u
mkdir -p gpurun_out/r2al
O=gpurun_out/r2al
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "default bench rc=$?"
python - <<PY
import json
d = json.loads(open("$O/bench_default.json").read().strip().splitlines()[-1])
print("value %.4g" % d["value"], "ms %.2f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], "frac %.3f" % d["roofline"]["frac"], "launches", d["gpu_launches"], {k: (v.get("value"), v.get("ms_per_step"), v.get("error")) for k, v in d["other_workloads"].items()})
PY
