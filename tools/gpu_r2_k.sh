#!/bin/bash
set -u
mkdir -p gpurun_out/r2l
O=gpurun_out/r2l
timeout 900 python -m pytest tests/test_equiv_gpu.py tests/test_plain_cnn_gpu.py tests/test_equiv_split_gpu.py -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench_ppo.json 2> $O/bench_ppo.err
timeout 600 python bench.py --workload equiv --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_equiv_fp32.json 2> $O/bench_equiv_fp32.err
timeout 600 python bench.py --workload equiv --precision bf16 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_equiv_bf16.json 2> $O/bench_equiv_bf16.err
timeout 600 python bench.py --workload cnn --steps 5 --warmup 3 > $O/bench_cnn_fp32.json 2> $O/bench_cnn_fp32.err
tail -4 $O/pytest_gpu.log; tail -2 $O/smoke.log
python - <<'PY'
import json
for f in ("ppo","equiv_fp32","equiv_bf16","cnn_fp32"):
    try:
        d=json.loads(open(f"gpurun_out/r2l/bench_{f}.json").read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"], d["gpu_launches"])
    except Exception as e: print(f,"ERR",e)
PY
if grep -q "rc=0" $O/pytest_gpu.log; then bash tools/gpu_r2_g.sh; else echo "tests failed: profiles skipped"; fi
