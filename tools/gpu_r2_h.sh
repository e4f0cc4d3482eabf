#!/bin/bash
set -u
mkdir -p gpurun_out/r2i
O=gpurun_out/r2i
timeout 1500 python -m pytest tests/test_equiv_split_gpu.py -q -s > $O/pytest_split.log 2>&1; echo "rc=$?" >> $O/pytest_split.log
timeout 1800 python -m pytest tests -m gpu -q --deselect tests/test_equiv_split_gpu.py > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 600 python bench.py --workload equiv --precision split --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_equiv_split.json 2> $O/bench_equiv_split.err
timeout 600 python bench.py --workload equiv --steps 5 --warmup 3 > $O/bench_equiv_fp32.json 2> $O/bench_equiv_fp32.err
timeout 600 python bench.py --workload equiv --precision bf16 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_equiv_bf16.json 2> $O/bench_equiv_bf16.err
timeout 600 python bench.py --workload cnn --steps 5 --warmup 3 > $O/bench_cnn_fp32.json 2> $O/bench_cnn_fp32.err
grep "rel \|device vs\|routing dec\|own-routing" $O/pytest_split.log | cut -c1-260; tail -3 $O/pytest_split.log; tail -4 $O/pytest_gpu.log
python - <<'PY'
import json
for f in ("equiv_fp32","equiv_split","equiv_bf16","cnn_fp32"):
    try:
        d=json.loads(open(f"gpurun_out/r2i/bench_{f}.json").read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"])
    except Exception as e: print(f,"ERR",e)
PY
