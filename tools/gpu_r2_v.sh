#!/bin/bash
# round 2, call V: hidden 128 end to end (training test, whole-iteration bench), default bench with the new other_workloads entry
mkdir -p gpurun_out/r2v
timeout 600 python -m pytest tests/test_train_gpu.py -x -q -m gpu > gpurun_out/r2v/pytest_train.log 2>&1; echo "train tests rc=$?"
tail -4 gpurun_out/r2v/pytest_train.log
for h in 128 256; do
  timeout 300 python bench.py --workload ppo --hidden_dim $h --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2v/bench_ppo_hidden$h.json 2> gpurun_out/r2v/bench_ppo_hidden$h.err; echo "bench hidden $h rc=$?"
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2v/bench_ppo_hidden$h.json").read().strip().splitlines()[-1])
print($h, "value %.4g" % d["value"], "ms %.2f" % d["ms_per_step"], d.get("phase_ms"), "launches", d.get("gpu_launches"), d["roofline"]["kernel"][:40], "frac %.3f" % d["roofline"]["frac"])
PY
done
timeout 600 python bench.py > gpurun_out/r2v/bench_default.json 2> gpurun_out/r2v/bench_default.err; echo "default bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/r2v/bench_default.json").read().strip().splitlines()[-1])
print("value %.4g" % d["value"], "ms %.2f" % d["ms_per_step"], {k: (v.get("value"), v.get("ms_per_step"), v.get("error")) for k, v in d["other_workloads"].items()})
PY
