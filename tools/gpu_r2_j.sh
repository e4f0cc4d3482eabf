#!/bin/bash
set -u
mkdir -p gpurun_out/r2j
O=gpurun_out/r2j
timeout 1500 python -m pytest tests/test_equiv_split_gpu.py tests/test_equiv_gpu.py tests/test_plain_cnn_gpu.py tests/test_guard_gpu.py tests/test_update_gpu.py -q -s > $O/pytest_sel.log 2>&1; echo "rc=$?" >> $O/pytest_sel.log
timeout 600 python bench.py --workload equiv --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_equiv_fp32.json 2> $O/bench_equiv_fp32.err
timeout 600 python bench.py --workload equiv --precision split --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_equiv_split.json 2> $O/bench_equiv_split.err
grep "gradient: rel\|2.1 M-sample\|device vs" $O/pytest_sel.log | cut -c1-220; tail -3 $O/pytest_sel.log
python - <<'PY'
import json
for f in ("equiv_fp32","equiv_split"):
    try:
        d=json.loads(open(f"gpurun_out/r2j/bench_{f}.json").read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"])
    except Exception as e: print(f,"ERR",e)
PY
if grep -q "rc=0" $O/pytest_sel.log; then bash tools/gpu_r2_g.sh; else echo "tests failed: profiles skipped"; fi
