"""GPU test of the tensor-core policy forward (tcgen05, TF32 operands, fp32 accumulate) against a
plain PyTorch fp32 evaluation of the same two MLPs (the reference's Agent networks,
control/ppo/ppo_model_gsde.py:37-52).  Tolerance: TF32 keeps 10 mantissa bits of each operand
(relative 5e-4 per product) and the hidden tanh is MUFU.TANH (abs error ~5e-4); the sums have 5 / 64
terms.  With PPO-initialised weights the outputs agree to < 2e-3 absolute; the stress case
inflates every weight by 0.3*randn (outputs up to +-4) and is held to 8e-3 * (1 + |ref|) (measured 3-5e-3)."""
import math

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def f16():
    import f16_model_b200 as m

    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    m.load_library()
    return m


def _ref(policy, obs):
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        m = policy.actor_mean(obs.double().float()).squeeze(-1)
        v = policy.critic(obs).squeeze(-1)
    torch.backends.cuda.matmul.allow_tf32 = prev
    return m, v


@pytest.mark.parametrize("precision", ["tf32", "fp16"])
@pytest.mark.parametrize("n", [1, 127, 128, 129, 4096, 65536 + 77, 148 * 512 * 3 + 5])
def test_policy_forward_matches_torch_fp32(f16, n, precision):
    """Both forward kernels -- policy_forward_kernel (TF32 operands) and policy_forward16_kernel (fp16 operands, the rollout
    kernel's four-groups-per-CTA schedule; the last size gives every group of every CTA several tiles and a ragged tail) --
    against torch fp32: the same bounds (fp16 and TF32 carry the same 11 significant bits)."""
    pol = f16.TensorCorePolicy(seed=3)
    # make the output layers non-trivial (the PPO init gives the actor head std 0.01)
    with torch.no_grad():
        for net in (pol.actor_mean, pol.critic):
            for p in net.parameters():
                p.add_(torch.randn_like(p) * 0.3)
    pol.sync_weights()
    g = torch.Generator(device="cuda").manual_seed(n)
    obs = torch.rand((n, 5), device="cuda", generator=g) * 2 - 1
    mean, value = pol.mean_value(obs, precision=precision)
    rm, rv = _ref(pol, obs)
    assert torch.isfinite(mean).all() and torch.isfinite(value).all()
    em = ((mean - rm).abs() / (1 + rm.abs())).max().item()
    ev = ((value - rv).abs() / (1 + rv.abs())).max().item()
    assert em < 8e-3 and ev < 8e-3, (em, ev)
    # PPO initialisation (orthogonal, actor head std 0.01, critic head std 1)
    pol0 = f16.TensorCorePolicy(seed=4)
    m0, v0 = pol0.mean_value(obs, precision=precision)
    r0m, r0v = _ref(pol0, obs)
    assert (m0 - r0m).abs().max().item() < 2e-3 and (v0 - r0v).abs().max().item() < 2e-3


@pytest.mark.parametrize("precision", ["tf32", "fp16"])
def test_policy_act_sampling_and_logprob(f16, precision):
    import functools

    pol = f16.TensorCorePolicy(seed=5, logstd_init=-0.7)
    n = 1 << 16
    obs = torch.rand((n, 5), device="cuda") * 2 - 1
    a_other, lp_other, _, mean_other = pol.act(obs, sample=True, step=11, want_mean=True, precision="fp16" if precision == "tf32" else "tf32")
    pol.act = functools.partial(pol.act, precision=precision)
    a, lp, v, mean = pol.act(obs, sample=True, step=11, want_mean=True)
    # the two kernels draw the SAME noise (same counter keys): identical log-probabilities, actions apart by the means' difference
    assert torch.equal(lp, lp_other)
    assert ((a - mean) - (a_other - mean_other)).abs().max().item() < 1e-6 and (mean - mean_other).abs().max().item() < 4e-3
    rm, rv = _ref(pol, obs)
    std = math.exp(-0.7)
    eps = (a - mean) / std
    assert abs(eps.mean().item()) < 0.02 and abs(eps.std().item() - 1.0) < 0.02          # N(0,1) noise
    ref_lp = -0.5 * eps ** 2 + 0.7 - 0.5 * math.log(2 * math.pi)
    assert (lp - ref_lp).abs().max().item() < 1e-4
    assert (mean - rm).abs().max().item() < 4e-3 and (v - rv).abs().max().item() < 4e-3
    # same (seed, step) -> same noise; different step -> different noise; sample=False -> the mean
    a2, _, _ = pol.act(obs, sample=True, step=11)
    a3, _, _ = pol.act(obs, sample=True, step=12)
    a4, lp4, _ = pol.act(obs, sample=False, step=11)
    assert torch.equal(a, a2) and not torch.equal(a, a3)
    assert torch.allclose(a4, mean) and torch.allclose(lp4, torch.full_like(lp4, 0.7 - 0.5 * math.log(2 * math.pi)))
    # the differentiable torch path used by the update agrees with the kernel's log-prob
    lp_t, ent, v_t = pol.evaluate(obs, a)
    assert (lp_t.detach() - lp).abs().max().item() < 2e-2
    assert (v_t.detach() - v).abs().max().item() < 4e-3


def test_policy_drives_the_env_on_device(f16):
    """Closed loop without a host round trip: obs -> tensor-core policy -> action -> env step."""
    cfg = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": True, "scenario": "combo"}
    env = f16.F16VecEnv(cfg, 8192, seed=1)
    pol = f16.TensorCorePolicy(seed=1)
    obs, _ = env.reset(seed=1)
    total = torch.zeros(8192, device="cuda")
    for k in range(50):
        a, lp, v = pol.act(obs, step=k)
        obs, r, d, _, _ = env.step(a)
        total += r
    assert torch.isfinite(total).all() and torch.isfinite(obs).all()
    # unit-variance exploration noise slams the stick now and then: a few envs trip the 50 deg/s limit
    # (-1000), the rest collect tracking reward
    assert (total > -900).float().mean().item() > 0.9 and (total > 0).float().mean().item() > 0.8


def _gae_reference(rewards, values, dones, next_value, next_done, gamma, lam):
    """The reference's Python loop (control/ppo/ppo_model_gsde.py:244-268), in torch on the device."""
    T = rewards.shape[0]
    adv = torch.zeros_like(rewards)
    lastgaelam = 0
    for t in reversed(range(T)):
        if t == T - 1:
            nextnonterminal, nextvalues = 1.0 - next_done, next_value
        else:
            nextnonterminal, nextvalues = 1.0 - dones[t + 1], values[t + 1]
        delta = rewards[t] + gamma * nextvalues * nextnonterminal - values[t]
        adv[t] = lastgaelam = delta + gamma * lam * nextnonterminal * lastgaelam
    return adv, adv + values


def test_gae_kernel_matches_reference_loop(f16):
    T, N = 257, 5000
    g = torch.Generator(device="cuda").manual_seed(0)
    rewards = torch.randn((T, N), device="cuda", generator=g)
    values = torch.randn((T, N), device="cuda", generator=g)
    dones = (torch.rand((T, N), device="cuda", generator=g) < 0.02).float()
    next_value = torch.randn(N, device="cuda", generator=g)
    next_done = (torch.rand(N, device="cuda", generator=g) < 0.02).float()
    done_after = torch.cat([dones[1:], next_done[None]]).to(torch.uint8)        # flag returned BY step t
    adv, ret = f16.gae(rewards, values, done_after, next_value, 0.99, 0.95)
    radv, rret = _gae_reference(rewards, values, dones, next_value, next_done, 0.99, 0.95)
    assert (adv - radv).abs().max().item() < 2e-4 and (ret - rret).abs().max().item() < 2e-4


def test_ppo_graph_update_equals_eager_update(f16):
    """The CUDA-graph replay of the optimiser step must reproduce the eager step (same minibatch order)."""
    cfg = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": True, "scenario": "combo"}
    out = []
    for use_graph in (True, False):
        envs = f16.F16VecEnv(cfg, 2048, seed=4)
        pol = f16.TensorCorePolicy(seed=4, logstd_init=-1.0)
        tr = f16.PPOTrainer(envs, pol, f16.PPOConfig(num_steps=32, num_minibatches=4, update_epochs=2, learning_rate=1e-3,
                                                      use_cuda_graph=use_graph))
        tr.train(num_updates=2)
        out.append([p.detach().clone() for p in pol.parameters()])
    for a, b in zip(*out):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6), (a - b).abs().max().item()


def test_ppo_updates_run_on_device(f16):
    """Two PPO updates (rollout -> GAE -> clipped update) on 4096 envs: finite losses, parameters move,
    the tensor-core weight blob follows the optimiser, and the value function starts explaining returns."""
    cfg = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": True, "scenario": "combo"}
    envs = f16.F16VecEnv(cfg, 4096, seed=1)
    pol = f16.TensorCorePolicy(seed=1, logstd_init=-1.0)
    trainer = f16.PPOTrainer(envs, pol, f16.PPOConfig(num_steps=64, num_minibatches=8, update_epochs=4, learning_rate=1e-3))
    before = [p.detach().clone() for p in pol.parameters()]
    blob_before = pol._blob.clone()
    hist = trainer.train(num_updates=2)
    assert len(hist) == 2 and all(np.isfinite(h["value_loss"]) and np.isfinite(h["policy_loss"]) for h in hist)
    assert trainer._fused_rollout and trainer._graph is None   # the rollout is ONE persistent kernel launch (f16_rollout), no graph needed
    assert trainer._update_graph not in (None, False)       # ... and so was the optimiser step
    assert any(not torch.equal(a, b) for a, b in zip(before, pol.parameters()))
    assert not torch.equal(blob_before, pol._blob)
    assert hist[-1]["global_step"] == 2 * 64 * 4096 and hist[-1]["rollout_sps"] > 1e6
    obs = trainer.obs[0]
    # graph replays draw fresh exploration noise (RNG keyed by the env's own counters)
    a_first = trainer.actions[5].clone()
    trainer.rollout()
    assert not torch.equal(a_first, trainer.actions[5])
    # kernel-side episode statistics seen through the trainer: a wild policy (std = e^1.5) crashes envs
    envs4 = f16.F16VecEnv(cfg, 4096, seed=2)
    pol4 = f16.TensorCorePolicy(seed=2, logstd_init=1.5)
    t4 = f16.PPOTrainer(envs4, pol4, f16.PPOConfig(num_steps=64))
    _, _, ep_ret, ep_cnt = t4.rollout()
    assert int(ep_cnt.item()) == int(t4.done_after.sum().item()) > 0
    assert ep_ret.item() < -900 * ep_cnt.item() * 0.5            # each crash carries the -1000 penalty
    # eager rollout == graph rollout, bit for bit
    envs2 = f16.F16VecEnv(cfg, 4096, seed=1)
    pol2 = f16.TensorCorePolicy(seed=1, logstd_init=-1.0)
    t_graph = f16.PPOTrainer(envs2, pol2, f16.PPOConfig(num_steps=16, fused_rollout=False))
    envs3 = f16.F16VecEnv(cfg, 4096, seed=1)
    pol3 = f16.TensorCorePolicy(seed=1, logstd_init=-1.0)
    t_eager = f16.PPOTrainer(envs3, pol3, f16.PPOConfig(num_steps=16, use_cuda_graph=False, fused_rollout=False))
    ag, rg, _, _ = t_graph.rollout()
    ae, re, _, _ = t_eager.rollout()
    assert t_graph._graph is not None and t_eager._graph is None   # the two-kernel path: replayed from a CUDA graph vs launched eagerly
    assert torch.equal(t_graph.actions, t_eager.actions) and torch.equal(t_graph.rewards, t_eager.rewards)
    assert torch.equal(ag, ae) and torch.equal(rg, re)
    m, v = pol.mean_value(obs)
    with torch.no_grad():
        assert (m - pol.actor_mean(obs).squeeze(-1)).abs().max().item() < 3e-3     # blob re-packed after the update


@pytest.mark.parametrize("precision", ["fp16", "tf32"])
@pytest.mark.parametrize("n,clip,norm_adv,clip_vloss", [(4096 + 37, 100.0, 1, 1), (128, 100.0, 1, 1), (65536, 0.2, 1, 1),
                                                       (2048 + 5, 100.0, 0, 0)])
def test_fused_ppo_gradient_matches_autograd(f16, n, clip, norm_adv, clip_vloss, precision):
    """f16_ppo_grad16 / f16_ppo_grad (forward + clipped-PPO loss + backward of both MLPs in one tcgen05 kernel, fp16 or TF32
    tensor-core operands) against torch autograd on the same minibatch.  With the clip range wide open the loss is smooth and the gradients agree to TF32 / MUFU.TANH level
    (< 2e-2 relative L2 per tensor, measured <= 7e-3); with clip_coef = 0.2 the rows whose ratio sits within the forward's
    ~2e-3 error of a clip boundary flip branch, which is noise of order sqrt(flipped rows) / n on a gradient whose own norm
    is of order 1 / sqrt(n) for this random data: held to cosine > 0.95, and the summed statistics to 1e-2."""
    from f16_model_b200 import _capi
    from f16_model_b200.policy import grad_views, pack_train_blob, pack_train_blob16

    torch.manual_seed(0)
    L = _capi.load_library()
    pol = f16.TensorCorePolicy(seed=3, logstd_init=-0.5)
    with torch.no_grad():
        for net in (pol.actor_mean, pol.critic):
            for p in net.parameters():
                p.add_(torch.randn_like(p) * 0.1)
    N = 200000
    obs = torch.rand((N, 5), device="cuda") * 2 - 1
    with torch.no_grad():
        mean0, val0 = pol.actor_mean(obs).squeeze(-1), pol.critic(obs).squeeze(-1)
    std = math.exp(-0.5)
    actions = mean0 + std * torch.randn(N, device="cuda")
    logprobs = -0.5 * ((actions - mean0) / std) ** 2 + 0.5 - 0.5 * math.log(2 * math.pi) + 0.3 * torch.randn(N, device="cuda")
    adv = torch.randn(N, device="cuda") * 2 + 0.3
    ret = val0 + torch.randn(N, device="cuda")
    val = val0 + 0.4 * torch.randn(N, device="cuda")
    mb = torch.randperm(N, device="cuda")[:n]
    vf, ent = 0.5, 0.01
    nl, entropy, nv = pol.evaluate(obs[mb], actions[mb])
    logratio = nl - logprobs[mb]
    ratio = logratio.exp()
    a = adv[mb]
    if norm_adv:
        a = (a - a.mean()) / (a.std() + 1e-8)
    pg = torch.max(-a * ratio, -a * torch.clamp(ratio, 1 - clip, 1 + clip)).mean()
    if clip_vloss:
        vc = val[mb] + torch.clamp(nv - val[mb], -clip, clip)
        vl = 0.5 * torch.max((nv - ret[mb]) ** 2, (vc - ret[mb]) ** 2).mean()
    else:
        vl = 0.5 * ((nv - ret[mb]) ** 2).mean()
    (pg - ent * entropy.mean() + vf * vl).backward()
    la = [m for m in pol.actor_mean if isinstance(m, torch.nn.Linear)]
    lc = [m for m in pol.critic if isinstance(m, torch.nn.Linear)]
    ref = {"w1": torch.stack([la[0].weight.grad, lc[0].weight.grad]), "b1": torch.stack([la[0].bias.grad, lc[0].bias.grad]),
           "w2": torch.stack([la[1].weight.grad, lc[1].weight.grad]), "b2": torch.stack([la[1].bias.grad, lc[1].bias.grad]),
           "w3": torch.stack([la[2].weight.grad.reshape(-1), lc[2].weight.grad.reshape(-1)]),
           "b3": torch.cat([la[2].bias.grad, lc[2].bias.grad]), "logstd": pol.actor_logstd.grad,
           "stats": torch.stack([(-logratio).sum(), ((ratio - 1) - logratio).sum(), ((ratio - 1).abs() > clip).float().sum(),
                                 vl * n, pg * n]).detach()}
    blob = (pack_train_blob16 if precision == "fp16" else pack_train_blob)(pol.actor_mean, pol.critic)
    g = torch.zeros(L.f16_ppo_grad_floats(), device="cuda")
    stats = torch.stack([adv[mb].mean(), adv[mb].std()])
    args = (n, blob.data_ptr(), obs.data_ptr(), actions.data_ptr(), logprobs.data_ptr(), adv.data_ptr(), ret.data_ptr(), val.data_ptr(),
            mb.data_ptr(), pol.actor_logstd.data_ptr(), stats.data_ptr(), g.data_ptr(), clip, vf, ent, norm_adv, clip_vloss)
    if precision == "fp16":
        # with scratch (per-CTA partials + reduction kernel: grad is overwritten) and without (atomics into the zeroed grad)
        g_atomic = torch.zeros_like(g)
        _capi.check(L.f16_ppo_grad16(*args[:11], g_atomic.data_ptr(), *args[12:], None, _capi.stream_ptr(obs.device)))
        scratch = torch.empty(L.f16_ppo_grad16_scratch_floats(0), device="cuda")
        g.fill_(123.0)
        _capi.check(L.f16_ppo_grad16(*args, scratch.data_ptr(), _capi.stream_ptr(obs.device)))
        assert torch.allclose(g, g_atomic, rtol=1e-4, atol=1e-7)
        # packed rows (f16_ppo_pack_rows: one 32-byte record + one (ret, val) pair per rollout row): the same gradient
        rows, rv = torch.empty((N, 8), device="cuda"), torch.empty((N, 2), device="cuda")
        _capi.check(L.f16_ppo_pack_rows(N, obs.data_ptr(), actions.data_ptr(), logprobs.data_ptr(), adv.data_ptr(), ret.data_ptr(), val.data_ptr(),
                                        rows.data_ptr(), rv.data_ptr(), _capi.stream_ptr(obs.device)))
        assert torch.equal(rows[:, :5], obs) and torch.equal(rows[:, 7], adv) and torch.equal(rv[:, 0], ret) and torch.equal(rv[:, 1], val)
        g_packed = torch.empty_like(g)
        _capi.check(L.f16_ppo_grad16(n, blob.data_ptr(), rows.data_ptr(), None, None, None, rv.data_ptr(), None, *args[8:11], g_packed.data_ptr(),
                                     *args[12:], scratch.data_ptr(), _capi.stream_ptr(obs.device)))
        Gp, Gu = grad_views(g_packed), grad_views(g)
        for k in ("w1", "b1", "w2", "b2", "w3"):          # tensor-core / fixed-order sums: bit for bit
            assert torch.equal(Gp[k], Gu[k]), k
        assert torch.allclose(g_packed, g, rtol=1e-5, atol=1e-8)   # b3, logstd and the statistics meet in shared-memory atomics
    else:
        _capi.check(L.f16_ppo_grad(*args, _capi.stream_ptr(obs.device)))
    G = grad_views(g)
    for k, r in ref.items():
        r, x = r.float().reshape(-1), G[k].reshape(-1)
        assert torch.isfinite(x).all()
        rel = ((x - r).norm() / (r.norm() + 1e-12)).item()
        if k == "stats":
            assert rel < 1e-2, (k, rel)
        elif clip > 10:
            assert rel < 2e-2, (k, rel)
        else:
            cos = torch.nn.functional.cosine_similarity(x, r, dim=0).item() if r.numel() > 1 else float(torch.sign(x[0] * r[0]))
            assert cos > 0.95, (k, cos, rel)


def test_fused_adam_matches_torch_clip_and_adam(f16):
    """f16_ppo_adam == clip_grad_norm_(max_norm) + torch.optim.Adam(amsgrad=True) on the same gradients, and the blobs it
    re-packs equal flat[perm]."""
    from f16_model_b200 import _capi

    L = _capi.load_library()
    torch.manual_seed(1)
    n = 9347
    p0 = torch.randn(n, device="cuda")
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=3e-4, amsgrad=True)
    p = p0.clone()
    m, v, vmax, step = (torch.zeros(n, device="cuda") for _ in range(3)), None, None, torch.zeros(1, device="cuda")
    m, v, vmax = list(m)
    lr = torch.tensor(3e-4, device="cuda")
    perm = torch.randint(0, n + 1, (5000,), device="cuda")
    blob_a, blob_b = torch.empty(5000, device="cuda"), torch.empty(3000, device="cuda")
    blob_h = torch.empty(4000, device="cuda", dtype=torch.float16)
    for it in range(6):
        g = torch.randn(n, device="cuda") * (0.001 if it % 2 else 1.0)       # below and above the clipping threshold
        ref.grad = g.clone()
        torch.nn.utils.clip_grad_norm_([ref], 0.5)
        opt.step()
        _capi.check(L.f16_ppo_adam(n, p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), vmax.data_ptr(), step.data_ptr(),
                                   lr.data_ptr(), 0.9, 0.999, 1e-8, 0.5, perm.data_ptr(), blob_a.data_ptr(), 5000,
                                   perm.data_ptr(), blob_b.data_ptr(), 3000, perm.data_ptr(), blob_h.data_ptr(), 4000,
                                   None, None, 0.0, 0, _capi.stream_ptr(p.device)))
        assert torch.allclose(p, ref.detach(), rtol=1e-5, atol=1e-7), (it, (p - ref.detach()).abs().max().item())
    assert step.item() == 6
    padded = torch.cat([torch.zeros(1, device="cuda"), p])
    assert torch.equal(blob_a, padded[perm]) and torch.equal(blob_b, padded[perm[:3000]])
    assert torch.equal(blob_h, padded[perm[:4000]].to(torch.float16))


def test_adv_stats_kernel_matches_torch(f16):
    from f16_model_b200 import _capi

    L = _capi.load_library()
    torch.manual_seed(2)
    adv = torch.randn(500000, device="cuda") * 3 + 0.7
    out = torch.zeros(2, device="cuda")
    scratch = torch.zeros(4, device="cuda", dtype=torch.float64)
    for n in (1, 2, 1000, 262144):
        mb = torch.randperm(500000, device="cuda")[:n]
        _capi.check(L.f16_ppo_adv_stats(n, 1, adv.data_ptr(), mb.data_ptr(), out.data_ptr(), scratch.data_ptr(), _capi.stream_ptr(adv.device)))
        a = adv[mb]
        ref = torch.stack([a.mean(), a.std() if n > 1 else torch.zeros((), device="cuda")])
        assert torch.allclose(out, ref, rtol=2e-6, atol=1e-6), (n, out, ref)
    assert scratch.abs().sum().item() == 0          # re-armed for the next call
    # a whole epoch in one launch: 7 minibatches of 5000 rows
    perm = torch.randperm(500000, device="cuda")[:35000]
    out7 = torch.zeros((7, 2), device="cuda")
    scratch7 = torch.zeros((7, 4), device="cuda", dtype=torch.float64)
    _capi.check(L.f16_ppo_adv_stats(5000, 7, adv.data_ptr(), perm.data_ptr(), out7.data_ptr(), scratch7.data_ptr(), _capi.stream_ptr(adv.device)))
    a7 = adv[perm].view(7, 5000)
    assert torch.allclose(out7, torch.stack([a7.mean(1), a7.std(1)], 1), rtol=2e-6, atol=1e-6)


@pytest.mark.parametrize("n", [1, 2, 1000, 65536, (1 << 20) + 17])
def test_randperm_kernel_is_a_permutation(f16, n):
    """f16_ppo_randperm: the epoch's minibatch order (np.random.permutation in ppo_model_gsde.py:296-300) as a keyed bijection of
    [0, n) -- every index exactly once, reproducible per (n, seed), different per seed, and mixed: every slice of the output
    (a minibatch) spreads over the whole range."""
    from f16_model_b200 import _capi
    L = _capi.load_library()
    outs = []
    for seed in (1, 1, 2):
        out = torch.full((n,), -1, dtype=torch.long, device="cuda")
        _capi.check(L.f16_ppo_randperm(n, seed, out.data_ptr(), _capi.stream_ptr(out.device)))
        assert torch.equal(torch.sort(out).values, torch.arange(n, device="cuda"))
        outs.append(out)
    assert torch.equal(outs[0], outs[1])
    if n >= 1000:
        assert not torch.equal(outs[0], outs[2])
        assert (outs[0] == torch.arange(n, device="cuda")).float().mean().item() < 0.01      # few fixed points
        if n >= 65536:
            for part in outs[0].float().chunk(8):                                            # each eighth looks uniform on [0, n)
                assert abs(part.mean().item() / n - 0.5) < 0.02 and abs(part.std().item() / n - 0.2887) < 0.01
        # neighbours in the output are not neighbours in the input
        assert (outs[0][1:] - outs[0][:-1]).abs().float().median().item() > n / 20


def test_fused_step16_equals_gradient_then_adam(f16):
    """f16_ppo_step16 (gradient kernel + ONE multi-CTA kernel: reduction, global-norm meeting, clip + Adam(amsgrad), blob scatter)
    against the two-call path f16_ppo_grad16 -> f16_ppo_adam on the same minibatches: identical gradients, parameters / Adam state equal
    up to the summation order of the gradient norm, and every blob word equal to a fresh packing of the parameters it ends with."""
    import ctypes as C

    from f16_model_b200 import _capi
    from f16_model_b200.policy import pack_policy_blob, pack_train_blob16

    cfg = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": False, "scenario": "combo"}
    trainers = []
    for fused in (True, False):
        envs = f16.F16VecEnv(cfg, 4096, seed=2, fast_math=True)
        pol = f16.TensorCorePolicy(seed=2, logstd_init=-0.5)
        tr = f16.PPOTrainer(envs, pol, f16.PPOConfig(num_steps=16, num_minibatches=4, update_epochs=3, learning_rate=3e-3, max_grad_norm=0.05,
                                                     fused_optimizer=fused, use_cuda_graph=fused))
        assert (tr._step16 is not None) == fused
        tr.train(num_updates=2)
        trainers.append(tr)
    a, b = trainers
    n = a._n_params
    assert float(a._adam[3].item()) == float(b._adam[3].item()) == 2 * 3 * 4            # optimiser steps taken
    # max_grad_norm = 0.05 clips every step: the global norm (a grid-wide sum in the fused kernel) really enters the update
    rel = lambda x, y: ((x - y).norm() / (y.norm() + 1e-12)).item()   # noqa: E731
    assert rel(a._pflat[:n], b._pflat[:n]) < 2e-5, rel(a._pflat[:n], b._pflat[:n])
    for k in range(3):
        assert rel(a._adam[k], b._adam[k]) < 1e-3
    # the scattered blobs ARE the packing of the fused trainer's own parameters
    pol = a.policy
    assert torch.equal(pol._blob, pack_policy_blob(pol.actor_mean, pol.critic))
    assert torch.equal(a._train_blob, pack_train_blob16(pol.actor_mean, pol.critic))
    assert torch.isfinite(a._diag).all() and abs(a._diag[5].item() - (pol.actor_logstd.item() + 1.4189385332046727)) < 1e-6


@pytest.mark.parametrize("n_envs,num_minibatches", [(4096, 4), (512, 8), (520, 8), (32768, 2)])
def test_fused_epoch16_equals_step16_per_minibatch(f16, n_envs, num_minibatches):
    """f16_ppo_epoch16 (ALL minibatch steps of an epoch in one persistent cooperative kernel, the optimiser step taken inside it between
    grid-wide meetings) against f16_ppo_step16 called once per minibatch on the same minibatches: the same gradients (same kernel code,
    same reduction order), parameters / Adam state equal up to the summation order of the gradient norm, the same diagnostics, and
    every blob word equal to a fresh packing of the final parameters.  Grids of 128 CTAs (two 64-parameter passes per CTA), 8 CTAs
    (a CTA owns more than 512 parameters; with 520 envs the minibatch is 1040 rows: a ragged last tile) and 148 CTAs (the production shape)."""
    from f16_model_b200.policy import pack_policy_blob, pack_train_blob16

    cfg = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": False, "scenario": "combo"}
    trainers, hist = [], []
    for epoch in (True, False):
        envs = f16.F16VecEnv(cfg, n_envs, seed=2, fast_math=True)
        pol = f16.TensorCorePolicy(seed=2, logstd_init=-0.5)
        tr = f16.PPOTrainer(envs, pol, f16.PPOConfig(num_steps=16, num_minibatches=num_minibatches, update_epochs=3, learning_rate=3e-3,
                                                     max_grad_norm=0.05, fused_epoch=epoch))
        hist.append(tr.train(num_updates=2))
        trainers.append(tr)
    a, b = trainers
    n = a._n_params
    assert float(a._adam[3].item()) == float(b._adam[3].item()) == 2 * 3 * num_minibatches
    rel = lambda x, y: ((x - y).norm() / (y.norm() + 1e-12)).item()   # noqa: E731
    assert rel(a._pflat[:n], b._pflat[:n]) < 5e-5, rel(a._pflat[:n], b._pflat[:n])   # up to 48 clipped steps: only the norm's summation order differs
    for k in range(3):
        assert rel(a._adam[k], b._adam[k]) < 1e-3
    assert torch.allclose(a._diag, b._diag, rtol=2e-3, atol=1e-5), (a._diag, b._diag)
    assert rel(a._g[:n], b._g[:n]) < 1e-2, rel(a._g[:n], b._g[:n])   # the last minibatch's gradient (parameters differ by ~1e-5: fp16 roundings flip)
    assert torch.allclose(a._g[n:], b._g[n:], rtol=2e-3, atol=1e-3)  # and its five summed statistics
    pol = a.policy
    assert torch.equal(pol._blob, pack_policy_blob(pol.actor_mean, pol.critic))
    assert torch.equal(a._train_blob, pack_train_blob16(pol.actor_mean, pol.critic))
    for k in ("value_loss", "approx_kl", "entropy"):
        assert abs(hist[0][-1][k] - hist[1][-1][k]) <= 2e-3 * max(1.0, abs(hist[1][-1][k])), (k, hist[0][-1][k], hist[1][-1][k])
