"""Probe: the fused PPO-gradient kernel (f16_ppo_grad) against torch autograd on the same minibatch."""
import os, sys, math, torch
sys.path.insert(0, os.getcwd())
import f16_model_b200 as f16
from f16_model_b200 import _capi
from f16_model_b200.policy import pack_train_blob, pack_train_blob16, grad_views

torch.manual_seed(0)
L = _capi.load_library()
pol = f16.TensorCorePolicy(seed=3, logstd_init=-0.5)
with torch.no_grad():
    for net in (pol.actor_mean, pol.critic):
        for p in net.parameters():
            p.add_(torch.randn_like(p) * 0.1)
N, n = 600000, int(sys.argv[1]) if len(sys.argv) > 1 else 4096 + 37
obs = torch.rand((N, 5), device="cuda") * 2 - 1
with torch.no_grad():
    mean0 = pol.actor_mean(obs).squeeze(-1)
    val0 = pol.critic(obs).squeeze(-1)
std = math.exp(-0.5)
actions = mean0 + std * torch.randn(N, device="cuda")
logprobs = -0.5 * ((actions - mean0) / std) ** 2 + 0.5 - 0.5 * math.log(2 * math.pi) + 0.3 * torch.randn(N, device="cuda")
adv = torch.randn(N, device="cuda") * 2 + 0.3
ret = val0 + torch.randn(N, device="cuda")
val = val0 + 0.4 * torch.randn(N, device="cuda")
mb = torch.randperm(N, device="cuda")[:n]
clip, vf, ent = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2, 0.5, 0.01

# ---- autograd reference (ppo.py _minibatch_step arithmetic)
for p in pol.parameters():
    p.grad = None
nl, entropy, nv = pol.evaluate(obs[mb], actions[mb])
logratio = nl - logprobs[mb]
ratio = logratio.exp()
a = adv[mb]
a = (a - a.mean()) / (a.std() + 1e-8)
pg = torch.max(-a * ratio, -a * torch.clamp(ratio, 1 - clip, 1 + clip)).mean()
vu = (nv - ret[mb]) ** 2
vc = val[mb] + torch.clamp(nv - val[mb], -clip, clip)
vl = 0.5 * torch.max(vu, (vc - ret[mb]) ** 2).mean()
loss = pg - ent * entropy.mean() + vf * vl
loss.backward()
la = [m for m in pol.actor_mean if isinstance(m, torch.nn.Linear)]
lc = [m for m in pol.critic if isinstance(m, torch.nn.Linear)]
ref = {"w1": torch.stack([la[0].weight.grad, lc[0].weight.grad]), "b1": torch.stack([la[0].bias.grad, lc[0].bias.grad]),
       "w2": torch.stack([la[1].weight.grad, lc[1].weight.grad]), "b2": torch.stack([la[1].bias.grad, lc[1].bias.grad]),
       "w3": torch.stack([la[2].weight.grad.reshape(-1), lc[2].weight.grad.reshape(-1)]),
       "b3": torch.cat([la[2].bias.grad, lc[2].bias.grad]), "logstd": pol.actor_logstd.grad,
       "stats": torch.stack([(-logratio).sum(), ((ratio - 1) - logratio).sum(), ((ratio - 1).abs() > clip).float().sum(), vl * n, pg * n]).detach()}

# ---- fused kernel
FP16 = os.environ.get("PPO_GRAD", "fp16") == "fp16"
blob = pack_train_blob16(pol.actor_mean, pol.critic) if FP16 else pack_train_blob(pol.actor_mean, pol.critic)
fn = L.f16_ppo_grad16 if FP16 else L.f16_ppo_grad
g = torch.zeros(L.f16_ppo_grad_floats(), device="cuda")
scratch = torch.empty(L.f16_ppo_grad16_scratch_floats(0), device="cuda")
am = adv[mb]
stats = torch.stack([am.mean(), am.std()])
def run():
    g.zero_()
    extra = (scratch.data_ptr(),) if FP16 else ()
    _capi.check(fn(n, blob.data_ptr(), obs.data_ptr(), actions.data_ptr(), logprobs.data_ptr(), adv.data_ptr(), ret.data_ptr(),
                               val.data_ptr(), mb.data_ptr(), pol.actor_logstd.data_ptr(), stats.data_ptr(), g.data_ptr(), clip, vf, ent, 1, 1,
                               *extra, _capi.stream_ptr(obs.device)))
run()
torch.cuda.synchronize()
G = grad_views(g)
worst = 0.0
for k in ref:
    r, x = ref[k].float().reshape(-1), G[k].reshape(-1)
    rel = ((x - r).norm() / (r.norm() + 1e-12)).item()
    cos = torch.nn.functional.cosine_similarity(x, r, dim=0).item() if r.numel() > 1 else 1.0
    worst = max(worst, rel)
    print(f"{k:7s} rel L2 err {rel:9.2e}  cos {cos:.6f}  |ref| {r.norm().item():.3e}")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3): run()
e0.record()
for _ in range(20): run()
e1.record(); torch.cuda.synchronize()
print("n =", n, " %.1f us per fused gradient (incl. zeroing)" % (e0.elapsed_time(e1) / 20 * 1e3), " worst rel err %.2e" % worst)
