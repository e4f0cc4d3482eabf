#!/bin/bash
set -u
mkdir -p gpurun_out/r2p
O=gpurun_out/r2p
for c in 8 16 32; do
  AUR_CONV_CHUNK=$c timeout 300 python bench.py --workload equiv --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_equiv_fp32_c$c.json 2> $O/err_c$c.txt
  AUR_CONV_CHUNK=$c timeout 300 python bench.py --workload equiv --precision split --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_equiv_split_c$c.json 2>> $O/err_c$c.txt
done
AUR_CONV_CHUNK=32 timeout 600 python -m pytest tests/test_equiv_split_gpu.py -q -s -k "conv_layer or full_update" > $O/pytest_c32.log 2>&1; echo "c32 tests rc=$?"
grep "conv Cin 1024\|device vs" $O/pytest_c32.log | cut -c1-200
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2p/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f.split("/")[-1], "%.1f ms" % d["ms_per_step"], d["clocks"]["sm_mhz"])
    except Exception as e: print(f, "ERR", e)
PY
