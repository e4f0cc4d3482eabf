#!/bin/bash
# 8-GPU box: scaling runs of the headline workload (weak) and of config E (1M envs, strong), wait statistics, DP tests on real peers
set -u
mkdir -p gpurun_out/r2f
O=gpurun_out/r2f
nvidia-smi -L > $O/gpus.txt 2>&1
run() { # n workload tag extra
  n=$1; w=$2; tag=$3; shift 3
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) \
    bench.py --gpus $n --workload $w --steps 20 --warmup 5 --no-cpu-baseline "$@" > $O/bench_${w}_n${n}_${tag}.json 2> $O/bench_${w}_n${n}_${tag}.err
  echo "$w n=$n $tag rc=$?"
}
timeout 600 python -m pytest tests/test_dp_peer_gpu.py -q > $O/pytest_dp.log 2>&1; echo "dp tests rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_ppo_n1_a.json 2> $O/bench_ppo_n1_a.err
run 2 ppo a; run 4 ppo a
for t in a b c d e; do run 8 ppo $t; done
AUR_DP_EXCHANGE=nccl run 8 ppo nccl
run 8 scale1m a --steps 10
run 4 scale1m a --steps 10
run 2 scale1m a --steps 10
run 8 pendulum a --steps 5 --warmup 3
tail -2 $O/pytest_dp.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2f/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.4g" % d["value"], "ms %.2f" % d["ms_per_step"], d["phase_ms"], (d.get("dp_wait") or {}).get("grad_wall_wait_us_per_minibatch_max_rank"))
    except Exception as e:
        print(f, "ERR", e)
PY
