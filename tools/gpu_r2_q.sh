#!/bin/bash
set -u
mkdir -p gpurun_out/r2q
O=gpurun_out/r2q
timeout 900 python -m pytest tests/test_equiv_split_gpu.py -q -s -k "config_d" > $O/pytest_cfgd.log 2>&1; echo "cfgd rc=$?"; grep "additivity\|passed\|failed\|Error" $O/pytest_cfgd.log | head -5
timeout 2400 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench_ppo.json 2> $O/bench_ppo.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err
python - <<'PY'
import json
for f in ("ppo","ref"):
    try:
        d=json.loads(open(f"gpurun_out/r2q/bench_{f}.json").read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d.get("roofline",{}).get("traffic"), d.get("cpu_baseline",{}).get("value"), d.get("e2e"))
    except Exception as e: print(f,"ERR",e)
PY
