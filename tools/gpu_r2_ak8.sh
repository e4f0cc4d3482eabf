#!/bin/bash
# round 2, call AK (8 GPUs): the hidden-128 iteration data-parallel over 8 GPUs
mkdir -p gpurun_out/r2ak
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --hidden_dim 128 --steps 5 --warmup 3 > gpurun_out/r2ak/bench_ppo_hidden128_n8.json 2> gpurun_out/r2ak/bench_ppo_hidden128_n8.err; echo "bench n=8 hidden 128 rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/r2ak/bench_ppo_hidden128_n8.json").read().strip().splitlines()[-1])
print("n_gpus", d["n_gpus"], "value %.4g" % d["value"], "ms %.2f" % d["ms_per_step"], d.get("phase_ms"))
PY
