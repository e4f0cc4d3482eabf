"""Times aur_ppo_update_grad for the wide policy shapes, layer-wise tensor-core path vs the shape-generic SIMT kernel, beside
the fused 64 x 2 kernel.  Usage: python tools/bench_wide.py [m]   -> one JSON line per (shape, path)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from aur_ppo_b200 import _lib, kernels

m = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
SHAPES = [(4, 2, 64, 2, False), (4, 2, 128, 2, False), (4, 2, 256, 2, False), (4, 2, 128, 3, False), (3, 1, 128, 2, True)]
if os.environ.get("WIDE_ONLY"):
    SHAPES = [(4, 2, int(os.environ["WIDE_ONLY"]), 2, False)]
L = _lib.lib()
for obs_dim, act_dim, H, NL, cont in SHAPES:
    for wide in ((1,) if (H == 64 or os.environ.get('WIDE_ONLY')) else (1, 0)):
        L.aur_ppo_update_set_wide(wide)
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
        L.aur_launch_count_reset()
        e0.record()
        n = 3
        for _ in range(n):
            up.grad(*bufs, idx)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        flops = 6.0 * P * m       # fwd + 2 x bwd, 2 flop per parameter and sample
        print(json.dumps({"shape": [obs_dim, act_dim, H, NL, cont], "path": "fused tc" if H == 64 else ("wide tc" if wide else "generic simt"),
                          "m": m, "ms": round(ms, 3), "samples_per_s": m / ms * 1e3, "fp32_equiv_tflops": flops / ms / 1e9,
                          "launches": L.aur_launch_count() // n, "workspace_MB": up.workspace.numel() * 4 / 2**20}), flush=True)
        del up
L.aur_ppo_update_set_wide(1)
