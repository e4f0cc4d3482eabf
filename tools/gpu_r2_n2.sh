#!/bin/bash
set -u
mkdir -p gpurun_out/r2n
O=gpurun_out/r2n
run() { n=$1; w=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) \
    bench.py --gpus $n --workload $w --no-cpu-baseline "$@" > $O/bench_${w}_n${n}.json 2> $O/bench_${w}_n${n}.err; echo "$w n=$n rc=$?"; }
timeout 600 python -m pytest tests/test_dp_peer_gpu.py -q > $O/pytest_dp.log 2>&1; echo "dp tests rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29777 tools/train_dp_smoke.py > $O/train_dp.log 2>&1; echo "train rc=$?"; tail -2 $O/train_dp.log
run 2 ppo --steps 20 --warmup 5
run 2 pendulum --steps 5 --warmup 3
run 2 scale1m --steps 5 --warmup 3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29778 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err; echo "ref rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2n/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f.split("/")[-1], "%.4g" % d["value"], d.get("ms_per_step"), d.get("dp_wait"), d.get("cpu_baseline", {}) and d["cpu_baseline"].get("cores"))
    except Exception as e: print(f, "ERR", e)
PY
