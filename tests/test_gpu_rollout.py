"""The closed-loop rollout kernel (csrc/f16_rollout.cu: policy MLPs on tcgen05 -> Gaussian head -> F16.step, K steps per
launch, state and observation resident on the SM) against the two-kernel path it replaces
(control/ppo/ppo_model_gsde.py:216-230: get_action_and_value, then envs.step, per step).

  * env side: the trajectories are BIT-identical to `F16VecEnv.step` fed with the fused kernel's own actions -- rewards,
    done flags, observations at every step, and every persistent buffer at the end, through terminations, autoresets and
    the 1000-step time-out;
  * policy side: at every step the fused kernel's action, log-probability and value agree with
    `policy_forward_kernel` evaluated on the same observation with the same RNG counters to <= 2e-3 (fp16 vs TF32
    tensor-core operands, both 11 significant bits, MUFU.TANH in both) -- and through it with torch fp32
    (tests/test_gpu_policy.py)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

CFG = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": False, "scenario": "combo"}


@pytest.fixture(scope="module")
def f16():
    import f16_model_b200 as m

    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device; the product has no CPU fallback")
    m.load_library()
    return m


def _buffers(K, n):
    z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device="cuda")     # noqa: E731
    return dict(obs_next=z(K, n, 5), actions=z(K, n), logprobs=z(K, n), values=z(K, n), rewards=z(K, n),
                dones=z(K, n, dt=torch.uint8), next_value=z(n))


def _spiced_env(f16, n, seed, fast):
    """Random-reset envs, a few of them started next to the limits so that terminations happen from the first steps on."""
    env = f16.F16VecEnv(CFG, n, seed=seed, fast_math=fast)
    st = env.state.clone()
    st[2, 5:40:5] = 0.86            # 49.3 deg/s: one pull away from the 50 deg/s limit
    st[1, 7:70:9] = 303.0           # just above the 300 m floor, descending
    st[3, 7:70:9] = -0.2
    env.set_state(st, ref_idx=env.ref_idx.clone())
    obs, _ = env.reset(keep_state=True, keep_reference=True)
    return env, obs.clone()


@pytest.mark.parametrize("fast,n,K,logstd", [(True, 2048 + 77, 1010, 0.0), (False, 640, 300, -0.5), (True, 65536, 40, -1.0)])
def test_fused_rollout_equals_two_kernel_path(f16, fast, n, K, logstd):
    pol = f16.TensorCorePolicy(seed=11, logstd_init=logstd)
    with torch.no_grad():           # a non-trivial actor head (the PPO init gives it std 0.01): actions saturate for part of the fleet
        pol.actor_mean[4].weight.mul_(30.0)
    pol.sync_weights()
    blob16 = f16.pack_train_blob16(pol.actor_mean, pol.critic)

    env_f, obs0 = _spiced_env(f16, n, 21, fast)
    env_r, obs0_r = _spiced_env(f16, n, 21, fast)
    assert torch.equal(obs0, obs0_r) and torch.equal(env_f.state, env_r.state)
    out = _buffers(K, n)
    f16.fused_rollout(env_f, blob16, pol.actor_logstd.detach(), obs0, **out, sample=True, policy_seed=pol.seed)
    torch.cuda.synchronize()

    obs = obs0
    worst_a = worst_v = worst_lp = 0.0
    n_done = 0
    for k in range(K):
        a_ref, lp_ref, v_ref = pol.act(obs, sample=True, env=env_r)            # policy_forward_kernel, the env's counters as RNG counters
        worst_a = max(worst_a, (out["actions"][k] - a_ref).abs().max().item())
        worst_v = max(worst_v, (out["values"][k] - v_ref).abs().max().item())
        worst_lp = max(worst_lp, (out["logprobs"][k] - lp_ref).abs().max().item())
        o, r, d, _, _ = env_r.step(out["actions"][k])                          # env_step2_kernel on the fused kernel's actions
        assert torch.equal(o, out["obs_next"][k]), k
        assert torch.equal(r, out["rewards"][k]) and torch.equal(d.view(torch.uint8), out["dones"][k]), k
        n_done += int(d.sum()) if (k % 16 == 15 or k == K - 1) else 0
        obs = out["obs_next"][k]
    for name in ("state", "episode_length", "total_return", "ref_idx", "episode_no", "episode_stats", "_final_obs", "_final_len", "_final_ret"):
        assert torch.equal(getattr(env_f, name), getattr(env_r, name)), name
    assert torch.equal(env_f._obs, env_r._obs)
    assert int(out["dones"].sum()) > (n if K > 1000 else 4), "the window must contain terminations"
    _, v_boot = pol.mean_value(out["obs_next"][K - 1])
    worst_v = max(worst_v, (out["next_value"] - v_boot).abs().max().item())
    print(f"fused vs two-kernel: |d action| {worst_a:.2e}, |d value| {worst_v:.2e}, |d logprob| {worst_lp:.2e}, episodes ended {int(out['dones'].sum())}")
    assert worst_a < 2e-3 and worst_v < 2e-3 and worst_lp < 1e-5


def test_fused_rollout_deterministic_actions_and_torch_reference(f16):
    """sample=0: action == mean; mean and value against plain torch fp32 on the recorded observations."""
    n, K = 1024 + 5, 24
    pol = f16.TensorCorePolicy(seed=3, logstd_init=-1.0)
    blob16 = f16.pack_train_blob16(pol.actor_mean, pol.critic)
    env = f16.F16VecEnv(CFG, n, seed=2, fast_math=True)
    obs0 = env.reset()[0].clone()
    out = _buffers(K, n)
    f16.fused_rollout(env, blob16, pol.actor_logstd.detach(), obs0, **out, sample=False, policy_seed=0)
    obs_all = torch.cat([obs0[None], out["obs_next"][:-1]])
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        m = pol.actor_mean(obs_all.reshape(-1, 5)).reshape(K, n)
        v = pol.critic(obs_all.reshape(-1, 5)).reshape(K, n)
    torch.backends.cuda.matmul.allow_tf32 = prev
    assert (out["actions"] - m).abs().max().item() < 2e-3 and (out["values"] - v).abs().max().item() < 2e-3
    lp = -(-1.0) - 0.5 * np.log(2 * np.pi)
    assert torch.allclose(out["logprobs"], torch.full_like(out["logprobs"], lp), atol=1e-6)


def test_trainer_uses_the_fused_rollout_and_learns(f16):
    """PPOTrainer with fused_rollout (default) on the deterministic combo reference: the mean step reward rises, and
    the first rollout's buffers equal those of a trainer running the two-kernel graph to the policy tolerance."""
    det = dict(CFG, determenistic_ref=True)
    stats = {}
    for fused in (True, False):
        envs = f16.F16VecEnv(det, 4096, seed=3, fast_math=True)
        pol = f16.TensorCorePolicy(seed=3, logstd_init=-1.0)
        tr = f16.PPOTrainer(envs, pol, f16.PPOConfig(num_steps=250, num_minibatches=16, update_epochs=5, learning_rate=1e-3, clip_coef=0.2,
                                                      anneal_lr=False, fused_rollout=fused))
        assert tr._fused_rollout is fused
        hist = tr.train(num_updates=12)
        stats[fused] = [h["mean_step_reward"] for h in hist]
    print("mean step reward per update, fused / two-kernel:", [round(x, 3) for x in stats[True]], [round(x, 3) for x in stats[False]])
    assert abs(stats[True][0] - stats[False][0]) < 0.02              # same policy, same seeds: the first rollouts agree
    assert stats[True][-1] > stats[True][0] + 0.15 and stats[False][-1] > stats[False][0] + 0.15
