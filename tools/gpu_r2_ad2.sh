#!/bin/bash
# round 2, call AD (2 GPUs): data-parallel run of the wide paths over real peer memory
mkdir -p gpurun_out/r2ad
timeout 600 python -m pytest tests/test_dp_peer_gpu.py -x -q -m gpu -k "128" > gpurun_out/r2ad/pytest_dp.log 2>&1; echo "dp tests rc=$?"
tail -3 gpurun_out/r2ad/pytest_dp.log
for h in 128 256; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --hidden_dim $h --steps 5 --warmup 3 > gpurun_out/r2ad/bench_ppo_hidden${h}_n2.json 2> gpurun_out/r2ad/bench_ppo_hidden${h}_n2.err; echo "bench n=2 hidden $h rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/r2ad/bench_ppo_hidden${h}_n2.json").read().strip().splitlines()[-1])
print($h, "n_gpus", d["n_gpus"], "value %.4g" % d["value"], "ms %.2f" % d["ms_per_step"], d.get("phase_ms"), d.get("dp_wait"))
PY
done
