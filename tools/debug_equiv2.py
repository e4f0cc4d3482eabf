import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, math
import torch.nn.functional as F
from aur_ppo_b200 import equiv
from oracle import equiv_ref as Q
B = 8
params = equiv.init_params(seed=5, scale=1.1)
for k in ("actor.head.psi_triv", "actor.head.psi_irrep", "critic.head2.w"): params[k].mul_(0.1)
g = torch.Generator().manual_seed(1)
obs = torch.rand(B, 1, 128, 128, generator=g) * 0.32; state = (torch.rand(B, generator=g) > 0.5).float()
action = torch.randn(B, 5, generator=g); adv, ret = torch.randn(B, generator=g), torch.randn(B, generator=g)
cpu = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in params.items()}
with torch.no_grad(): lp0, _, v0 = Q.evaluate(cpu, state, obs, action)
oldlp = lp0 + 0.15 * torch.randn(B, generator=g); vold = v0 + 0.3 * torch.randn(B, generator=g)
# reference with intermediates
x = Q.cat_obs(state, obs)
acts_a, acts_c = [], []
fa = Q.encoder_forward(cpu, "actor", x, acts_a, quant=True); fc = Q.encoder_forward(cpu, "critic", x, acts_c, quant=True)
for t in acts_a + acts_c + [fa, fc]: t.retain_grad()
W = torch.cat([Q.expand_regular_to_irrep1(cpu["actor.head.psi_irrep"]), Q.expand_regular_to_trivial(cpu["actor.head.psi_triv"])], 0)
W = Q.bf16_ste(W)
out = fa @ W.T + torch.cat([torch.zeros(2), cpu["actor.head.bias_triv"]]); out.retain_grad()
mean = torch.cat((out[:, 2:3], out[:, 0:2], out[:, 3:5]), 1); log_std = torch.clamp(out[:, 5:], -20, 2)
std = log_std.exp()
lp = (-((action - mean) ** 2) / (2 * std ** 2) - std.log() - math.log(math.sqrt(2 * math.pi))).sum(1)
ent = (0.5 + 0.5 * math.log(2 * math.pi) + std.log()).sum(1)
W1 = Q.expand_regular_to_regular(cpu["critic.head1.psi"]).reshape(512, 512)
W1 = Q.bf16_ste(W1)
hpre = fc @ W1.T + Q.expand_bias_regular(cpu["critic.head1.bias"]); hpre.retain_grad()
h = F.relu(hpre); pooled = h.reshape(B, -1, 4).max(2).values
val = (pooled @ cpu["critic.head2.w"].T + cpu["critic.head2.bias"]).reshape(-1)
ratio = (lp - oldlp).exp(); a_n = (adv - adv.mean()) / (adv.std() + 1e-8)
pl = torch.max(-a_n * ratio, -a_n * torch.clamp(ratio, 0.8, 1.2)).mean()
vl = 0.5 * torch.max((val - ret) ** 2, (vold + torch.clamp(val - vold, -0.2, 0.2) - ret) ** 2).mean() * 0.5
loss = pl - 0.01 * ent.mean() + vl
loss.backward()
m = equiv.EquivActorCritic(params, B)
dev = lambda t: t.cuda().contiguous()
# instrument: capture intermediates by monkeypatching _encoder_backward
caps = {}
orig = m._encoder_backward
def patched(net, st, ob, dfeat):
    caps[net + ".dfeat"] = dfeat.clone(); return orig(net, st, ob, dfeat)
m._encoder_backward = patched
m.loss_and_grads(dev(state), dev(obs), dev(action), dev(oldlp), dev(adv), dev(ret), dev(vold))
def rel(a, b): return float((a - b).norm() / (b.norm() + 1e-20))
print("feat_a rel", rel(m.enc["actor"].feat.float().cpu(), fa.detach()), "feat_c rel", rel(m.enc["critic"].feat.float().cpu(), fc.detach()))
print("dfeat_a (pre-mask) vs ref dfeat*? :")
ref_dfa = fa.grad; ref_dfc = fc.grad
mask_a = (fa.detach() > 0).float(); mask_c = (fc.detach() > 0).float()
print("  dfa masked rel", rel(caps["actor.dfeat"].cpu() * mask_a, ref_dfa * mask_a), " dfc masked rel", rel(caps["critic.dfeat"].cpu() * mask_c, ref_dfc * mask_c))
print("  mask mismatch a", float(((m.enc['actor'].feat.float().cpu() > 0) != (fa.detach() > 0)).float().mean()))
for i, (t, name) in enumerate(zip(acts_a, ["a1", "a2", "a3", "a4", "a5", "a6"])):
    got = m.enc["actor"].a[i].float().cpu()
    off = 1 if i < 4 else 0
    H = t.shape[2]
    got = got[:, off:off + H, off:off + H, :].permute(0, 3, 1, 2)
    print(name, "fwd rel", rel(got, t.detach()))
a_out_g, c_pre_g, d_a_out_g, d_c_h_g = m._last_head
print("c_pre rel", rel(c_pre_g.cpu() + Q.expand_bias_regular(cpu["critic.head1.bias"]).detach(), hpre.detach()))
print("d_c_h rel", rel(d_c_h_g.float().cpu(), hpre.grad), "nnz mine", int((d_c_h_g != 0).sum()), "nnz ref", int((hpre.grad != 0).sum()))
mism = ((d_c_h_g.float().cpu() != 0) != (hpre.grad != 0))
print("nonzero pattern mismatches", int(mism.sum()))
print("d_a_out rel", rel(d_a_out_g.float().cpu()[:, :10], out.grad))
hp = hpre.detach().reshape(B, 128, 4)
top2 = hp.topk(2, dim=2).values
gap = ((top2[:, :, 0] - top2[:, :, 1]) / top2[:, :, 0].abs().clamp_min(1e-9))
act = top2[:, :, 0] > 0
print("active fields", int(act.sum()), "rel gap quantiles", torch.quantile(gap[act], torch.tensor([0.01, 0.05, 0.25, 0.5])))
cm = c_pre_g.cpu() + Q.expand_bias_regular(cpu["critic.head1.bias"]).detach()
print("max |c_pre - ref|", float((cm - hpre.detach()).abs().max()), "feat_c max diff", float((m.enc["critic"].feat.float().cpu() - fc.detach()).abs().max()))
print("value mine", m.value.cpu()[:4], "ref", val.detach()[:4])
sys.exit(0)
# ---- verify every tc GEMM call against torch matmul of the same operands
print("---- GEMM self-check")
orig_gemm = equiv.tc_gemm_bf16
def checked(a, b):
    c = orig_gemm(a, b)
    ref = a.float() @ b.float().T
    err = float((c - ref).abs().max()); sc = float(ref.abs().max())
    print("  gemm", tuple(a.shape), tuple(b.shape), "maxerr", err, "scale", sc, "BAD" if err > 1e-2 * sc + 1e-6 else "")
    return c
equiv.tc_gemm_bf16 = checked
m2 = equiv.EquivActorCritic(params, B)
m2.loss_and_grads(dev(state), dev(obs), dev(action), dev(oldlp), dev(adv), dev(ret), dev(vold))
