"""Short program for ncu: one equivariant (or, EQUIV_PLAIN=1, plain CNN) update at a small batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aur_ppo_b200 import equiv, plain_cnn
B = int(os.environ.get("EQUIV_B", "256"))
PLAIN = os.environ.get("EQUIV_PLAIN", "0") == "1"
PREC = os.environ.get("EQUIV_PRECISION", "split")
if PLAIN:
    params = plain_cnn.init_params(seed=0)
else:
    params = equiv.init_params(seed=0)
    for k in ("actor.head.psi_triv", "actor.head.psi_irrep", "critic.head2.w"):
        params[k].mul_(0.1)
g = torch.Generator(device="cuda").manual_seed(0)
obs = torch.rand(B, 1, 128, 128, generator=g, device="cuda") * 0.32
state = (torch.rand(B, generator=g, device="cuda") > 0.5).float()
action = torch.randn(B, 5, generator=g, device="cuda")
adv, ret, vold = (torch.randn(B, generator=g, device="cuda") for _ in range(3))
model = (plain_cnn.PlainActorCritic if PLAIN else equiv.EquivActorCritic)(params, B, precision=PREC)
for _ in range(2):
    model.update(state, obs, action, torch.full((B,), -7.0, device="cuda"), adv, ret, vold)
torch.cuda.synchronize()
print("equiv profile target done")
