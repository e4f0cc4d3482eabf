#!/bin/bash
# round 2, call X: whole GPU suite on the binary with the wide paths + ncu --set full of the wide update kernels (hidden 128)
set -u
mkdir -p gpurun_out/r2x
O=gpurun_out/r2x
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
WIDE_ONLY=128 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"wide_|tc_gemm_tn" --launch-skip 60 -c 12 -o $O/wide_full python tools/bench_wide.py > $O/ncu_wide.log 2>&1; echo "wide full rc=$?"
python tools/ncu_summary.py $O/wide_full.ncu-rep "layer-wise update kernels, hidden 128 (round 2)" > $O/wide_full_summary.md 2> $O/wide_full_summary.err
rm -f $O/wide_full.ncu-rep
grep -c "^## " $O/wide_full_summary.md
for h in 256; do
  timeout 300 python bench.py --workload ppo --hidden_dim $h --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_ppo_hidden$h.json 2> $O/bench_ppo_hidden$h.err; echo "bench hidden $h rc=$?"
done
python - <<PY
import json
d = json.loads(open("$O/bench_ppo_hidden256.json").read().strip().splitlines()[-1])
print(256, "value %.4g" % d["value"], "ms %.2f" % d["ms_per_step"], d.get("phase_ms"), "launches", d.get("gpu_launches"))
PY
