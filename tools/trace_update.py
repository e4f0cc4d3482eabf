"""Per-phase cycle counts of one CTA of the traced update kernel (debug build libaurppo_trace.so)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aur_ppo_b200 import kernels, _lib
from tools.microbench import _policy
desc, flat = _policy(False)
B, m = 8388608, 2097152
g = torch.Generator(device="cuda").manual_seed(3)
dbuf = [torch.randn(B, 4, generator=g, device="cuda") * 0.5, torch.randint(0, 2, (B,), generator=g, device="cuda").float(),
        -0.7 + 0.1 * torch.randn(B, generator=g, device="cuda"), torch.randn(B, generator=g, device="cuda"),
        torch.randn(B, generator=g, device="cuda"), torch.randn(B, generator=g, device="cuda")]
idx = torch.randperm(B, generator=g, device="cuda")[:m].to(torch.int32)
up = kernels.Updater(desc, flat.clone())
rec = kernels.pack_records(*dbuf)
for _ in range(3):
    up.step(*dbuf, idx, lr=2.5e-4, records=rec)
torch.cuda.synchronize()
out = (ctypes.c_ulonglong * 32)()
_lib.lib().aur_debug_trace(out)
n = out[20]
names = {0: "P8'+loop edge -> before wait WG", 1: "wait WG", 3: "dz1/h1/aux stores + sync", 4: "issue? + gather + e loads", 5: "wait FWD",
         6: "z ld + h2", 8: "head + sync + loss + sync + dout", 9: "wait AUX", 10: "dz2 + store + sync", 11: "(issue)", 12: "dW3 passes", 13: "h1 ahead"}
tot = sum(out[i] for i in range(14))
print("tiles traced", n, "cycles per tile", tot / max(n, 1))
for i in range(14):
    if out[i]:
        print(f"{i:2d} {out[i]/n:9.1f} cyc  {100*out[i]/tot:5.1f}%  {names.get(i,'')}")
