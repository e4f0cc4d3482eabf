#!/bin/bash
# round 2, call AF: ncu --set full of the remaining wide kernels (256-wide update GEMMs, 128-wide rollout / value kernels)
set -u
mkdir -p gpurun_out/r2af
O=gpurun_out/r2af
WIDE_ONLY=256 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"skinny_gemm|tc_gemm_tn" --launch-skip 20 -c 4 -o $O/w256 python tools/bench_wide.py > $O/ncu_w256.log 2>&1; echo "w256 rc=$?"
python tools/ncu_summary.py $O/w256.ncu-rep "256-wide update GEMMs (round 2)" > $O/w256_summary.md 2> $O/w256_summary.err; rm -f $O/w256.ncu-rep
cat > /tmp/roll128.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from aur_ppo_b200 import envs as denv, kernels
N, T, H = 65536, 128, 128
desc = kernels.policy_desc(4, 2, H, 2, False)
P = kernels.policy_param_count(desc)
flat = (torch.rand(P, device="cuda") - 0.5) * (2.0 / H ** 0.5)
env = denv.DeviceVecEnv("CartPole-v1", N)
env.reset(list(range(N)))
buf = kernels.RolloutBuffers(T, N, 4, (), "cuda")
for i in range(3):
    kernels.rollout(env, desc, flat, buf, seed=1, step0=i * T)
torch.cuda.synchronize()
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rollout_tc_kernel|critic_values_tc" --launch-skip 3 -c 3 -o $O/r128 python /tmp/roll128.py > $O/ncu_r128.log 2>&1; echo "r128 rc=$?"
python tools/ncu_summary.py $O/r128.ncu-rep "128-wide rollout / value kernels (round 2)" > $O/r128_summary.md 2> $O/r128_summary.err; rm -f $O/r128.ncu-rep
grep -c "^## " $O/w256_summary.md $O/r128_summary.md
