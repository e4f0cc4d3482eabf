"""GPU debug: per-tap error of the layer-0 weight gradient (tensor-core kernel vs SIMT kernel, bf16 and split) against a float64
autograd on the device's own routing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from aur_ppo_b200 import _lib, kernels
from oracle import equiv_ref as Q

def run(planes, B=5):
    g = torch.Generator().manual_seed(11)
    psi = torch.randn(16, 2, 3, 3, generator=g) * 0.3
    bias = 0.1 * torch.randn(16, generator=g)
    obs = torch.rand(B, 1, 128, 128, generator=g) * 0.32
    state = (torch.rand(B, generator=g) > 0.5).float()
    da1 = torch.randn(B, 64, 64, 64, generator=g) * 0.1
    if planes == 1:
        da1 = da1.bfloat16().float()
    out = torch.zeros(planes, B, 66, 66, 64, dtype=torch.bfloat16, device="cuda")
    arg = torch.zeros(B, 64, 64, 64, dtype=torch.uint8, device="cuda")
    ws = torch.zeros(64 * 18 + 64, device="cuda")
    dpsi, dbias = torch.zeros(16, 2, 3, 3, device="cuda"), torch.zeros(16, device="cuda")
    gpl = kernels.split_planes(da1.cuda()) if planes == 2 else da1.cuda().bfloat16().unsqueeze(0).contiguous()
    with kernels.tc_precision(planes):
        kernels.equiv_conv0(obs.cuda(), state.cuda(), psi.cuda(), bias.cuda(), out, arg)
        rc = _lib.lib().aur_equiv_conv0_wgrad(obs.cuda().data_ptr(), state.cuda().data_ptr(), gpl.data_ptr(), out.data_ptr(),
                                              arg.data_ptr(), B, ws.data_ptr(), dpsi.data_ptr(), dbias.data_ptr(), None)
    _lib.check(rc, "wgrad")
    torch.cuda.synchronize()
    x = Q.cat_obs(state, obs).double()
    pd, bd = psi.double().requires_grad_(True), bias.double().requires_grad_(True)
    z = F.conv2d(x, Q.expand_trivial_to_regular(pd), Q.expand_bias_regular(bd), padding=1)
    a = arg.permute(0, 3, 1, 2).long().cpu()
    pos = out[0, :, 1:65, 1:65, :].permute(0, 3, 1, 2).float().cpu() > 0
    y = Q._windows(z).gather(-1, a.unsqueeze(-1)).squeeze(-1) * pos.double()
    y.backward(da1.double().permute(0, 3, 1, 2))
    err = (dpsi.cpu().double() - pd.grad)
    print(f"planes={planes} impl={os.environ.get('AUR_CONV0_WGRAD','tc')}: rel {float(err.norm()/pd.grad.norm()):.2e}  bias rel "
          f"{float((dbias.cpu().double()-bd.grad).norm()/bd.grad.norm()):.2e}")
    print("  per (ci, tap) rms error / rms grad:", (err.pow(2).mean(0).sqrt() / pd.grad.pow(2).mean(0).sqrt()).numpy().round(5))
    print("  per field o:", (err.pow(2).mean((1, 2, 3)).sqrt() / pd.grad.pow(2).mean((1, 2, 3)).sqrt()).numpy().round(5))

run(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
