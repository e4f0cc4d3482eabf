"""Times aur_ppo_update_grad for a list of policy shapes (the shape-generic kernel; impl 3 forces it on the headline shape).
Usage: python tools/bench_generic.py [m]   -> one JSON line per shape."""
import json
import sys

import torch

from aur_ppo_b200 import _lib, kernels

m = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
SHAPES = [(4, 2, 64, 2, False, 1), (4, 2, 64, 2, False, 3), (4, 2, 64, 4, False, 3), (4, 2, 128, 2, False, 3),
          (4, 2, 256, 2, False, 3), (6, 3, 64, 2, False, 3), (3, 1, 64, 2, True, 3), (4, 2, 32, 2, False, 3)]
L = _lib.lib()
for obs_dim, act_dim, H, NL, cont, impl in SHAPES:
    L.aur_ppo_update_set_impl(impl)
    desc = kernels.policy_desc(obs_dim, act_dim, H, NL, cont)
    P = kernels.policy_param_count(desc)
    g = torch.Generator(device="cuda").manual_seed(0)
    params = (torch.rand(P, device="cuda", generator=g) - 0.5) * 0.2
    up = kernels.Updater(desc, params)
    B = m
    bufs = [torch.randn(B, obs_dim, device="cuda"), torch.randn(B, act_dim, device="cuda") if cont else
            torch.randint(0, act_dim, (B,), device="cuda").float(), -0.7 + 0.1 * torch.randn(B, device="cuda"),
            torch.randn(B, device="cuda"), torch.randn(B, device="cuda"), torch.randn(B, device="cuda")]
    idx = torch.randperm(B, device="cuda").to(torch.int32)
    for _ in range(2):
        up.grad(*bufs, idx)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 3
    for _ in range(n):
        up.grad(*bufs, idx)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    flops = 6.0 * P * m       # fwd + 2 x bwd, 2 flop per parameter and sample
    print(json.dumps({"shape": [obs_dim, act_dim, H, NL, cont], "impl": impl, "m": m, "ms": ms, "samples_per_s": m / ms * 1e3,
                      "fp32_tflops": flops / ms / 1e9}))
