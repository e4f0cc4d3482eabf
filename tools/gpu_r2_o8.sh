#!/bin/bash
set -u
mkdir -p gpurun_out/r2o
O=gpurun_out/r2o
run() { n=$1; w=$2; tag=$3; shift 3
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) \
    bench.py --gpus $n --workload $w --steps 20 --warmup 5 --no-cpu-baseline "$@" > $O/bench_${w}_n${n}_${tag}.json 2> $O/bench_${w}_n${n}_${tag}.err; echo "$w n=$n $tag rc=$?"; }
run 8 ppo a; run 8 ppo b; run 4 ppo a; run 8 pendulum a --steps 5 --warmup 3
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2o/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); w = d.get("dp_wait") or {}
        print(f.split("/")[-1], "%.4g" % d["value"], "%.2f ms" % d["ms_per_step"], d["phase_ms"], {k: v for k, v in w.items() if k != "what"})
    except Exception as e: print(f, "ERR", e)
PY
