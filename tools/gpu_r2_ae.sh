#!/bin/bash
# round 2, call AE: whole GPU suite + default bench + hidden 128 / 256 benches on the final binary
set -u
mkdir -p gpurun_out/r2ae
O=gpurun_out/r2ae
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"
for h in 128 256; do
  timeout 300 python bench.py --workload ppo --hidden_dim $h --steps 5 --warmup 3 --no-extras > $O/bench_ppo_hidden$h.json 2> $O/bench_ppo_hidden$h.err; echo "bench hidden $h rc=$?"
done
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "default bench rc=$?"
python - <<PY
import json
for h in (128, 256):
    d = json.loads(open("$O/bench_ppo_hidden%d.json" % h).read().strip().splitlines()[-1])
    print(h, "value %.4g" % d["value"], "ms %.2f" % d["ms_per_step"], d.get("phase_ms"), "cpu", d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None)
d = json.loads(open("$O/bench_default.json").read().strip().splitlines()[-1])
print("value %.4g" % d["value"], "ms %.2f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], "frac %.3f" % d["roofline"]["frac"], {k: (v.get("value"), v.get("ms_per_step"), v.get("error")) for k, v in d["other_workloads"].items()})
PY
