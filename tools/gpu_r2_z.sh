#!/bin/bash
# round 2, call Z: facade test, launch list of the 256-wide update and of the 256-wide rollout
mkdir -p gpurun_out/r2z
timeout 600 python -m pytest tests/test_equiv_gpu.py -x -q -m gpu -k "facade" > gpurun_out/r2z/pytest_facade.log 2>&1; echo "facade rc=$?"
tail -3 gpurun_out/r2z/pytest_facade.log
WIDE_ONLY=256 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 320 --csv --log-file gpurun_out/r2z/launches_wide256.csv python tools/bench_wide.py > gpurun_out/r2z/ncu_wide.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/r2z/launches_wide256.csv 2>/dev/null | grep "aur::\|tc::" | head -12
cat > /tmp/roll256.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from aur_ppo_b200 import envs as denv, kernels
N, T, H = 65536, 128, 256
desc = kernels.policy_desc(4, 2, H, 2, False)
P = kernels.policy_param_count(desc)
flat = (torch.rand(P, device="cuda") - 0.5) * (2.0 / H ** 0.5)
env = denv.DeviceVecEnv("CartPole-v1", N)
env.reset(list(range(N)))
buf = kernels.RolloutBuffers(T, N, 4, (), "cuda")
for i in range(2):
    kernels.rollout(env, desc, flat, buf, seed=1, step0=i * T)
torch.cuda.synchronize()
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 700 -c 700 --csv --log-file gpurun_out/r2z/launches_roll256.csv python /tmp/roll256.py > gpurun_out/r2z/ncu_roll.log 2>&1; echo "ncu roll rc=$?"
python tools/summarize_launches.py gpurun_out/r2z/launches_roll256.csv 2>/dev/null | grep "aur::\|tc::" | head -12
