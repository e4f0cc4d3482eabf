"""Per-kernel timing on one GPU (CUDA events, L2 flushed between iterations)."""
import argparse
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aur_ppo_b200 import kernels

PEAK = 6450.9
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, iters=20, warmup=5, flush=True):
    scratch = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if flush else None
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        if flush:
            scratch.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def bench_gae(args):
    for T, N in [(128, 65536), (256, 65536), (128, 131072), (512, 131072), (2048, 131072), (128, 1048576), (128, 4096)]:
        rew = torch.rand(T, N, device="cuda"); val = torch.randn(T, N, device="cuda")
        term = (torch.rand(T, N, device="cuda") < 1 / 200).float()
        nv = torch.randn(N, device="cuda"); nd = torch.zeros(N, device="cuda")
        out = (torch.empty_like(rew), torch.empty_like(rew))
        med, best = timeit(lambda: kernels.gae(rew, val, term, nv, nd, 0.99, 0.95, True, out))
        byt = 20 * T * N + 8 * N
        print(json.dumps({"kernel": "gae", "T": T, "N": N, "us_median": med * 1e6, "us_best": best * 1e6,
                          "GBps_median": byt / med / 1e9, "frac_of_measured_peak": byt / med / 1e9 / PEAK}))
        del rew, val, term, out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=["gae"])
    a = ap.parse_args()
    for w in a.which:
        globals()["bench_" + w](a)
