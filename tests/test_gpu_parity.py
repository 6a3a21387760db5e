"""GPU parity tests (B200 box, `pytest -m gpu`): the CUDA path, called through the C ABI,
against (a) the golden vectors produced by the unmodified reference and (b) the CPU oracle
on the same seeded inputs.

Bars (BASELINE.json north_star): fp64 trajectories within 1e-10 relative, bit-exact
termination flags and step counts; fp32 within 1e-4 relative after 1000 steps.  'Relative'
is |d| / max(|ref|, field scale) with the floors in tests/workloads.py."""
import os

import numpy as np
import pytest

from workloads import FIELD_SCALE, FP32_FIELD_SCALE, config2_inputs, fp32_storage_sensitivity, rel_err

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

FP64_RTOL = 1e-10
FP32_RTOL = 1e-4


@pytest.fixture(scope="module")
def f16():
    import f16_model_b200 as m

    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device; the product has no CPU fallback")
    m.load_library()
    return m


def dev(a, dtype=torch.float64):
    return torch.tensor(np.ascontiguousarray(a), dtype=dtype, device="cuda")


CFG = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": True, "scenario": "combo"}


# ------------------------------------------------------------------ data level
def test_coefficients_fp64_vs_reference(f16, golden, orc):
    from f16_model_b200.model import aero_coeffs

    x = golden["coef_in"]
    out = aero_coeffs(*[dev(x[:, j]) for j in range(4)]).cpu().numpy().T
    assert np.max(np.abs(out - golden["coef_out"])) < 1e-13          # coefficients are O(1)
    rng = np.random.default_rng(7)
    n = 200_000
    al, fi = rng.uniform(-0.8, 2.0, n), rng.uniform(-0.7, 0.7, n)
    wz, V = rng.uniform(-1.1, 1.1, n), rng.uniform(60, 400, n)
    out = aero_coeffs(dev(al), dev(fi), dev(wz), dev(V)).cpu().numpy().T
    assert np.max(np.abs(out - orc.coeffs(al, fi, wz, V))) < 1e-12


def test_coefficients_fp32(f16, golden):
    from f16_model_b200.model import aero_coeffs

    x = golden["coef_in"]
    out = aero_coeffs(*[dev(x[:, j], torch.float32) for j in range(4)]).cpu().numpy().T
    assert np.max(np.abs(out - golden["coef_out"])) < 2e-5


def test_thrust_vs_reference(f16, golden, orc):
    from f16_model_b200.model import thrust

    x = golden["thrust_in"]
    out = thrust(dev(x[:, 0]), dev(x[:, 1]), dev(x[:, 2])).cpu().numpy()
    assert np.max(rel_err(out, golden["thrust_out"], 1e3)) < 1e-13
    out32 = thrust(*[dev(x[:, j], torch.float32) for j in range(3)]).cpu().numpy()
    assert np.max(rel_err(out32, golden["thrust_out"], 1e3)) < 1e-5
    rng = np.random.default_rng(8)
    H, M, Pa = rng.uniform(-300, 32000, 100_000), rng.uniform(0, 1.7, 100_000), rng.uniform(0, 130, 100_000)
    out = thrust(dev(H), dev(M), dev(Pa)).cpu().numpy()
    assert np.max(rel_err(out, orc.thrust(H, M, Pa), 1e3)) < 1e-12


def test_atmosphere_vs_reference(f16, golden):
    from f16_model_b200.model import atmosphere

    out = atmosphere(dev(golden["isa_in"])).cpu().numpy().T
    assert np.max(np.abs(out / golden["isa_out"] - 1)) < 1e-14
    out32 = atmosphere(dev(golden["isa_in"], torch.float32)).cpu().numpy().T
    assert np.max(np.abs(out32 / golden["isa_out"] - 1)) < 3e-6


def test_atmosphere_against_published_table(f16):
    """Device ISA vs the published standard-atmosphere table (same table as tests/test_oracle_golden.py)."""
    from f16_model_b200.model import atmosphere
    from test_oracle_golden import ISA_TABLE

    h = np.array([r[0] for r in ISA_TABLE], dtype=float)
    for dtype, tol in ((torch.float64, 6e-5), (torch.float32, 7e-5)):
        out = atmosphere(dev(h, dtype)).double().cpu().numpy()
        assert np.max(np.abs(out[0] / np.array([r[1] for r in ISA_TABLE]) - 1)) < tol
        assert np.max(np.abs(out[1] / np.array([r[2] for r in ISA_TABLE]) - 1)) < 3e-5


def test_atmosphere_layer_table_consistent_on_device(f16):
    """The device ISA's layer table (above 11 km: tabulated base pressures) against the primary constants:
    pressure continuous across every layer boundary to the table's six digits (tests/test_oracle_golden.py)."""
    from f16_model_b200.model import atmosphere
    from test_oracle_golden import isa_pressure_jumps

    jumps = isa_pressure_jumps(lambda h: atmosphere(dev(h)).cpu().numpy())
    assert np.max(np.abs(jumps)) < 1e-5, jumps


# ------------------------------------------------------------------ model level
def test_rhs_fp64_vs_reference(f16, golden, orc):
    x, u, ref = golden["rhs_x"], golden["rhs_u"], golden["rhs_dx"]
    dx = f16.rhs(dev(x.T), dev(u.T)).cpu().numpy().T
    scale = np.maximum(np.abs(ref).max(axis=0), 1e-3)
    assert np.max(np.abs(dx - ref) / np.maximum(np.abs(ref), scale)) < 1e-12
    # large random batch against the oracle
    rng = np.random.default_rng(9)
    n = 100_000
    X = np.stack([rng.uniform(0, 5000, n), rng.uniform(320, 29000, n), rng.uniform(-0.9, 0.9, n),
                  rng.uniform(-0.8, 1.2, n), rng.uniform(80, 320, n), rng.uniform(-0.45, 1.7, n),
                  rng.uniform(-0.5, 0.5, n), rng.uniform(-1.3, 1.3, n), rng.uniform(0, 100, n)], 1)
    U = np.stack([rng.uniform(-0.5, 0.5, n), rng.uniform(-0.1, 1.1, n)], 1)
    ref = orc.rhs(X, U)
    dx = f16.rhs(dev(X.T), dev(U.T)).cpu().numpy().T
    scale = np.maximum(np.abs(ref).max(axis=0), 1e-3) * 1e-3
    assert np.max(np.abs(dx - ref) / np.maximum(np.abs(ref), scale)) < 1e-10


def test_model_step_matches_oracle(f16, orc):
    """K x F16model.step with a held control == the oracle's rollout without env logic."""
    x0, _ = config2_inputs(512, seed=5)
    u = np.stack([np.full(512, -0.05), np.zeros(512)])
    x = dev(x0.T)
    f16.model_step(x, dev(u), 0.01, steps=200)
    acts = np.full((200, 512), (-0.05 - (-0.4363323152065277)) / 0.8726646304130554 * 2 - 1)
    ref = orc.rollout(x0, acts, np.zeros((1, 1000)), np.zeros(512, dtype=np.int32), want_obs=False)
    live = ref["done"][-1] == 0       # the bare model has no termination; compare envs that never tripped it
    err = rel_err(x.cpu().numpy().T[live], ref["final_state"][live], FIELD_SCALE)
    assert err.max() < 1e-9           # action->command conversion rounds once more than the direct command


# ------------------------------------------------------------------ env level: golden trajectories
def _golden_env(f16, golden, dtype, autoreset=False, **kw):
    n = golden["traj_x0"].shape[0]
    env = f16.F16VecEnv(CFG, n, dtype=dtype, autoreset=autoreset, reference_bank=golden["traj_ref"], **kw)
    env.set_state(golden["traj_x0"].T, ref_idx=np.arange(n))
    return env


def test_golden_trajectories_fp64_step_by_step(f16, golden):
    """1000 x env.step against the unmodified reference: states, obs, reward within 1e-10;
    done flags and step counts bit-exact (incl. config 1 = column 0)."""
    env = _golden_env(f16, golden, torch.float64)
    acts = dev(golden["traj_actions"])
    T, n = golden["traj_actions"].shape
    S, O, R, D = [], [], [], []
    for k in range(T):
        obs, r, d, trunc, info = env.step(acts[k])
        S.append(env.state.T.clone()); O.append(obs.clone()); R.append(r.clone()); D.append(d.clone())
        assert not bool(trunc.any())
    S, O = torch.stack(S).cpu().numpy(), torch.stack(O).cpu().numpy()
    R, D = torch.stack(R).cpu().numpy(), torch.stack(D).cpu().numpy().astype(np.uint8)
    assert np.array_equal(D, golden["traj_done"])
    assert np.array_equal(env.episode_length.cpu().numpy(), golden["traj_ep_len"])
    assert rel_err(S, golden["traj_states"], FIELD_SCALE).max() < FP64_RTOL
    live = np.concatenate([np.ones((1, n), bool), golden["traj_done"][:-1] == 0])   # reward is 0 once frozen
    assert rel_err(R[live], golden["traj_reward"][live], 1.0).max() < FP64_RTOL
    assert np.abs(O - golden["traj_obs"]).max() < FP64_RTOL
    assert rel_err(env.total_return.cpu().numpy(), golden["traj_return"], 1.0).max() < FP64_RTOL


def test_golden_trajectories_fp64_k_steps_per_launch(f16, golden):
    """Same episodes with K = 125 steps per launch, observations emitted every step."""
    env = _golden_env(f16, golden, torch.float64)
    acts = dev(golden["traj_actions"])
    obs_all, rew_all, done_all = [], [], []
    for c in range(8):
        o, r, d = env.step_k(acts[c * 125:(c + 1) * 125], obs_every_step=True)
        obs_all.append(o.clone()); rew_all.append(r); done_all.append(d)
    O, D = torch.cat(obs_all).cpu().numpy(), torch.cat(done_all).cpu().numpy()
    assert np.array_equal(D, golden["traj_done"])
    assert np.array_equal(env.episode_length.cpu().numpy(), golden["traj_ep_len"])
    assert np.abs(O - golden["traj_obs"]).max() < FP64_RTOL
    assert rel_err(env.state.T.cpu().numpy(), golden["traj_states"][-1], FIELD_SCALE).max() < FP64_RTOL
    assert rel_err(env.total_return.cpu().numpy(), golden["traj_return"], 1.0).max() < FP64_RTOL


@pytest.mark.parametrize("fast", [False, True])
def test_golden_trajectories_fp32(f16, golden, fast):
    """fp32 mode: within 1e-4 relative of the reference after 1000 steps (all envs that ran
    the full episode), termination steps of the hard terminations unchanged."""
    env = _golden_env(f16, golden, torch.float32, fast_math=fast)
    acts = dev(golden["traj_actions"], torch.float32)
    _, r, d = env.step_k(acts)
    ep = env.episode_length.cpu().numpy()
    assert np.array_equal(ep, golden["traj_ep_len"])
    err = rel_err(env.state.T.cpu().numpy(), golden["traj_states"][-1], FP32_FIELD_SCALE)
    assert err.max() < FP32_RTOL, err.max(axis=0)
    assert rel_err(env.total_return.cpu().numpy(), golden["traj_return"], 1.0).max() < FP32_RTOL


def test_single_env_facade_config1(f16, golden):
    """BASELINE config 1 through the reference-shaped single-env API (NumPy in/out)."""
    x0, u0 = f16.get_trimmed_state_control()
    cfg = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": True, "scenario": "pure_step",
           "init_state": x0, "init_control": u0}
    env = f16.F16(cfg)
    obs0, info0 = env.reset()
    assert info0 == {} and np.abs(obs0 - golden["traj_obs0"][0]).max() < 1e-15
    with pytest.raises(TypeError):
        env.step([0.0])
    total = 0.0
    for k in range(1000):
        obs, r, done, trunc, info = env.step(np.array([golden["traj_actions"][k, 0]]))
        total += r
        assert trunc is False and done == bool(golden["traj_done"][k, 0])
        if k in (0, 499, 999):
            assert rel_err(info["state"].to_array(), golden["traj_states"][k, 0], FIELD_SCALE).max() < FP64_RTOL
            assert np.abs(obs - golden["traj_obs"][k, 0]).max() < FP64_RTOL
    assert info["episode_length"] == 1000 and info["clock"] == 10.0
    assert abs(total - golden["traj_return"][0]) < 1e-8 and abs(info["total_return"] - total) < 1e-9


def test_single_env_reset_seed_follows_reference_draws(f16, golden):
    """F16.reset(seed) reseeds the process-global RNGs like the reference (env/env.py:155-158): the random
    initial state and the random reference signal are the reference's own draws for that seed."""
    cfg = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": False, "scenario": "cos"}
    env = f16.F16(cfg)
    for k, seed in enumerate([1, 2, 3]):
        obs, _ = env.reset(seed=seed)
        x0 = golden["random_state_out"][k]
        assert np.array_equal(env.init_state.to_array(), x0)
        assert np.array_equal(env.ref_signal.theta_ref, golden[f"ref_rnd_cos_{seed}"])
        w150 = np.radians(150)
        ref0 = env.ref_signal.wz_ref[0]
        expect = np.array([x0[1] / 35000, (x0[2] + w150) / (2 * w150), x0[4] / 500, ref0, 0.5 * (ref0 - x0[2])])
        assert np.abs(obs - expect).max() < 1e-15
        assert env.episode_length == 0 and not env.done


# ------------------------------------------------------------------ BASELINE config 2
@pytest.fixture(scope="module")
def config2(orc):
    x0, acts = config2_inputs(4096)
    bank = __import__("f16_model_b200").build_reference_bank(0.01, 10, None, False)   # 144 signals
    ref_idx = (np.arange(4096) * 7) % bank.shape[0]
    ref = orc.rollout(x0, acts, bank, ref_idx, want_states=True, want_obs=True)
    ref["fp32_storage_sensitivity"] = np.max(fp32_storage_sensitivity(orc, x0, acts, bank, ref_idx, ref), axis=0)   # NaN-propagating
    return x0, acts, bank, ref_idx.astype(np.int32), ref


def test_config2_fp64_4096_envs_1000_steps(f16, config2):
    """4096 envs x 1000 steps, fp64, K = 50 per launch, per-step obs: every state at every
    50th step, every obs/reward/done at every step, vs the oracle."""
    x0, acts, bank, ref_idx, ref = config2
    env = f16.F16VecEnv(CFG, 4096, dtype=torch.float64, autoreset=False, reference_bank=bank)
    env.set_state(x0.T, ref_idx=ref_idx)
    a = dev(acts)
    worst_s = worst_o = worst_r = 0.0
    for c in range(20):
        o, r, d = env.step_k(a[c * 50:(c + 1) * 50], obs_every_step=True)
        sl = slice(c * 50, (c + 1) * 50)
        assert np.array_equal(d.cpu().numpy(), ref["done"][sl])                  # bit-exact termination flags
        worst_o = max(worst_o, np.abs(o.cpu().numpy() - ref["obs"][sl]).max())
        worst_r = max(worst_r, rel_err(r.cpu().numpy(), ref["reward"][sl], 1.0).max())
        worst_s = max(worst_s, rel_err(env.state.T.cpu().numpy(), ref["states"][c * 50 + 49], FIELD_SCALE).max())
    assert np.array_equal(env.episode_length.cpu().numpy(), ref["ep_len"])       # bit-exact step counts
    n_term = int((ref["ep_len"] < 1000).sum())
    assert n_term > 300, "workload must exercise early termination"
    assert worst_s < FP64_RTOL and worst_o < FP64_RTOL and worst_r < FP64_RTOL, (worst_s, worst_o, worst_r)
    assert rel_err(env.total_return.cpu().numpy(), ref["total_return"], 1.0).max() < FP64_RTOL


@pytest.mark.parametrize("fast", [False, True])
def test_config2_fp32_drift_bound(f16, config2, fast):
    """fp32 throughput mode on the config-2 inputs after 1000 steps, envs that flew the whole episode.
    STATED BOUND (north_star; DESIGN.md 'fp32 mode'): relative error <= 1e-4 after 1000 steps, with the per-field floors of
    tests/workloads.py.  It is asserted for EVERY env whose trajectory is not rounding-sensitive, where sensitivity is measured on
    the fp64 model itself (`fp32_storage_sensitivity`: the oracle's exact fp64 arithmetic with nothing but the state rounded to
    fp32 between steps, plus at most a half-ulp of jitter per step -- less than any fp32 implementation adds -- deviates by less
    than 5e-5).  The few sensitive envs (about ten of 3544: full-scale random stick with alpha pinned at the -20 deg table edge;
    env 1226 reaches 8.8e-5 from state storage alone and 3.4e-4 with the jitter; tests/test_fp32_sensitivity.py) are held to four
    times their own fp64-model deviation.  Early terminations may shift by a step near a
    threshold, so step counts must agree for > 99.5 % of envs."""
    x0, acts, bank, ref_idx, ref = config2
    env = f16.F16VecEnv(CFG, 4096, dtype=torch.float32, autoreset=False, reference_bank=bank, fast_math=fast)
    env.set_state(x0.T, ref_idx=ref_idx)
    env.step_k(dev(acts, torch.float32))
    ep = env.episode_length.cpu().numpy()
    sens = ref["fp32_storage_sensitivity"]
    full = (ref["ep_len"] == 1000) & (ep == 1000) & np.isfinite(sens)
    assert abs(int((ep < 1000).sum()) - int((ref["ep_len"] < 1000).sum())) <= 20
    assert np.mean(ep == ref["ep_len"]) > 0.995
    err = rel_err(env.state.T.cpu().numpy()[full], ref["final_state"][full], FP32_FIELD_SCALE)
    per_env, s_env = err.max(axis=1), sens[full]
    robust = s_env < 5e-5
    print("fp32 fast=%s: max %.2e, 99.9%% %.2e, median %.2e, envs > 1e-4: %d of %d; rounding-sensitive envs: %d (worst fp64-storage deviation %.2e)" %
          (fast, per_env.max(), np.quantile(per_env, 0.999), np.median(per_env), int((per_env > 1e-4).sum()), per_env.size,
           int((~robust).sum()), s_env.max()))
    assert robust.sum() >= per_env.size - 16
    assert per_env[robust].max() <= FP32_RTOL                                   # the stated bound, every robust env
    assert np.all(per_env[~robust] <= np.maximum(FP32_RTOL, 4.0 * s_env[~robust]))
    assert np.quantile(per_env, 0.999) < FP32_RTOL


# ------------------------------------------------------------------ structural properties at full size
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_k_invariance_bitwise(f16, dtype):
    """K steps in one launch == K launches of one step, bit for bit (state in registers
    changes nothing)."""
    n = 1 << 16
    envs = [f16.F16VecEnv(CFG, n, dtype=dtype, autoreset=True, seed=3) for _ in range(2)]
    g = torch.Generator(device="cuda").manual_seed(0)
    acts = torch.rand((16, n), device="cuda", dtype=dtype, generator=g) * 2 - 1
    _, r_k, d_k = envs[0].step_k(acts)
    for k in range(16):
        _, r, d, _, _ = envs[1].step(acts[k])
        assert torch.equal(r, r_k[k]) and torch.equal(d.view(torch.uint8), d_k[k])
    assert torch.equal(envs[0].state, envs[1].state)
    assert torch.equal(envs[0].episode_length, envs[1].episode_length)


@pytest.mark.parametrize("fast", [False, True])
def test_two_envs_per_thread_kernel_equals_one_env_kernel(f16, fast):
    """The fp32 autoreset path runs the two-envs-per-thread kernel on full 512-env tiles (and the one-env
    kernel on the remainder); without autoreset everything runs the one-env kernel, which is the one held to
    the reference above.  Same arithmetic -> bit-identical states, rewards and observations as long as no
    episode ends."""
    n = 70_000 + 37                      # full tiles + a ragged remainder
    a_env = f16.F16VecEnv(CFG, n, dtype=torch.float32, autoreset=True, seed=8, fast_math=fast)
    b_env = f16.F16VecEnv(CFG, n, dtype=torch.float32, autoreset=False, seed=8, fast_math=fast)
    assert torch.equal(a_env.state, b_env.state)
    g = torch.Generator(device="cuda").manual_seed(3)
    acts = (torch.rand((48, n), device="cuda", generator=g) - 0.5) * 0.2          # gentle stick: nobody terminates
    for k in range(8):                                                             # K = 1 launches
        oa, ra, da, _, _ = a_env.step(acts[k])
        ob, rb, db, _, _ = b_env.step(acts[k])
        assert torch.equal(oa, ob) and torch.equal(ra, rb) and not bool(da.any()) and not bool(db.any())
    oa, ra, da = a_env.step_k(acts[8:])                                            # K = 40 in one launch
    ob, rb, db = b_env.step_k(acts[8:])
    assert torch.equal(ra, rb) and torch.equal(oa, ob) and int(da.sum()) == 0
    assert torch.equal(a_env.state, b_env.state) and torch.equal(a_env.total_return, b_env.total_return)
    assert torch.equal(a_env.episode_length, b_env.episode_length)


def test_sharding_invariance_full_size(f16):
    """BASELINE config 3/5 size: 2^20 envs as one batch == two shards of 2^19 keyed by global env
    id (what two GPUs would hold): bitwise equal states, rewards, dones after resets + 64 steps."""
    from f16_model_b200.sharding import make_sharded_env

    n = 1 << 20
    cfg = dict(CFG, determenistic_ref=False)
    whole = f16.F16VecEnv(cfg, n, dtype=torch.float32, autoreset=True, seed=11, bank_size=64)
    parts = [make_sharded_env(cfg, n, r, 2, "cuda", dtype=torch.float32, autoreset=True, seed=11, bank_size=64)
             for r in range(2)]
    g = torch.Generator(device="cuda").manual_seed(1)
    acts = torch.rand((64, n), device="cuda", generator=g) * 2 - 1
    acts[:, ::5] = 1.0                       # a fifth of the envs slam the stick -> resets happen inside the window
    _, r, d = whole.step_k(acts)
    assert int(d.sum()) > 1000
    h = n // 2
    for rk, p in enumerate(parts):
        _, rp, dp = p.step_k(acts[:, rk * h:(rk + 1) * h].contiguous())
        assert torch.equal(rp, r[:, rk * h:(rk + 1) * h]) and torch.equal(dp, d[:, rk * h:(rk + 1) * h])
        assert torch.equal(p.state, whole.state[:, rk * h:(rk + 1) * h])
        assert torch.equal(p.ref_idx, whole.ref_idx[rk * h:(rk + 1) * h])


def test_full_size_batch_fp64_subset_vs_oracle(f16, orc):
    """BASELINE config 3/5 size in the parity mode: 2^20 envs x 120 steps in fp64 (K = 40 per launch);
    a strided subset of 2048 envs is re-flown by the oracle from the same inputs -- the result for an env
    must not depend on the batch it sits in."""
    n, T = 1 << 20, 120
    rng = np.random.default_rng(99)
    d5 = np.radians(5)
    al = rng.uniform(-d5, d5, n)
    x0 = np.zeros((n, 9))
    x0[:, 1], x0[:, 2], x0[:, 3] = rng.uniform(2000, 4000, n), rng.uniform(-d5, d5, n), al
    x0[:, 4], x0[:, 5] = rng.uniform(90, 250, n), al
    bank = __import__("f16_model_b200").build_reference_bank(0.01, 10, None, False)
    ref_idx = rng.integers(0, bank.shape[0], n).astype(np.int32)
    g = torch.Generator(device="cuda").manual_seed(5)
    acts = torch.rand((T, n), device="cuda", dtype=torch.float64, generator=g) * 2 - 1
    acts[:, ::3] = 1.0                                   # a third of the fleet pulls full aft stick -> terminations
    env = f16.F16VecEnv(CFG, n, dtype=torch.float64, autoreset=False, reference_bank=bank)
    env.set_state(x0.T, ref_idx=ref_idx)
    dones = []
    for c in range(3):
        _, _, d = env.step_k(acts[c * 40:(c + 1) * 40])
        dones.append(d)
    sub = np.arange(0, n, 512)
    ref = orc.rollout(x0[sub], acts[:, sub].cpu().numpy(), bank, ref_idx[sub], want_obs=False)
    st = env.state.T.cpu().numpy()[sub]
    assert np.array_equal(torch.cat(dones).cpu().numpy()[:, sub], ref["done"])
    assert np.array_equal(env.episode_length.cpu().numpy()[sub], ref["ep_len"])
    assert int((ref["ep_len"] < T).sum()) > 300
    assert rel_err(st, ref["final_state"], FIELD_SCALE).max() < FP64_RTOL
    assert rel_err(env.total_return.cpu().numpy()[sub], ref["total_return"], 1.0).max() < FP64_RTOL


def test_eight_million_envs_sharding_invariance(f16):
    """BASELINE config 5 size (8 Mi envs, here on one GPU): one batch == four shards keyed by global env id."""
    from f16_model_b200.sharding import make_sharded_env

    n = 1 << 23
    cfg = dict(CFG, determenistic_ref=False)
    whole = f16.F16VecEnv(cfg, n, dtype=torch.float32, autoreset=True, seed=21, bank_size=32, fast_math=True)
    g = torch.Generator(device="cuda").manual_seed(2)
    acts = torch.rand((4, n), device="cuda", generator=g) * 2 - 1
    _, r, d = whole.step_k(acts)
    q = n // 4
    for rk in (0, 3):
        part = make_sharded_env(cfg, n, rk, 4, "cuda", dtype=torch.float32, autoreset=True, seed=21, bank_size=32, fast_math=True)
        _, rp, dp = part.step_k(acts[:, rk * q:(rk + 1) * q].contiguous())
        assert torch.equal(rp, r[:, rk * q:(rk + 1) * q]) and torch.equal(dp, d[:, rk * q:(rk + 1) * q])
        assert torch.equal(part.state, whole.state[:, rk * q:(rk + 1) * q])
        del part


@pytest.mark.parametrize("dt,tn,scenario", [(0.01, 5, "pure_step"), (0.02, 20, "step"), (0.005, 10, "cos")])
def test_other_step_sizes_and_horizons(f16, orc, dt, tn, scenario):
    """config["dt"] / config["tn"] other than 0.01 / 10 (test/env_test.py uses tn=5): the horizon
    tn/dt - 1 and the reference length follow, trajectories still match the oracle."""
    cfg = dict(CFG, dt=dt, tn=tn, scenario=scenario)
    T = int(round(tn / dt))
    x0, acts = config2_inputs(256, seed=11, T=1000)
    acts = np.tile(acts, (3, 1))[:T]
    bank = __import__("f16_model_b200").build_reference_bank(dt, tn, scenario, True)
    assert bank.shape == (1, T)
    env = f16.F16VecEnv(cfg, 256, dtype=torch.float64, autoreset=False)
    env.set_state(x0.T)
    _, r, d = env.step_k(dev(acts))
    ref = orc.rollout(x0, acts, bank, np.zeros(256, dtype=np.int32), dt=dt, horizon=T - 1, want_obs=False)
    assert np.array_equal(d.cpu().numpy(), ref["done"]) and np.array_equal(env.episode_length.cpu().numpy(), ref["ep_len"])
    assert ref["ep_len"].max() == T                       # the time-out fires on step tn/dt
    assert rel_err(env.state.T.cpu().numpy(), ref["final_state"], FIELD_SCALE).max() < FP64_RTOL


def test_determinism_and_seed_sensitivity(f16):
    a = f16.F16VecEnv(CFG, 4096, seed=5)
    b = f16.F16VecEnv(CFG, 4096, seed=5)
    c = f16.F16VecEnv(CFG, 4096, seed=6)
    assert torch.equal(a.state, b.state) and not torch.equal(a.state, c.state)
    s = a.state.double().cpu().numpy()
    d5 = np.radians(5)
    assert s[1].min() >= 2000 and s[1].max() <= 4000 and s[4].min() >= 90 and s[4].max() <= 250
    assert np.abs(s[2]).max() <= d5 and np.abs(s[5]).max() <= d5 and np.array_equal(s[3], s[5])
    assert np.all(s[[0, 6, 7, 8]] == 0)
    # uniformity (distribution-equivalent to get_random_state): means within 4 sigma
    assert abs(s[1].mean() - 3000) < 4 * 2000 / np.sqrt(12 * 4096)
    assert abs(s[4].mean() - 170) < 4 * 160 / np.sqrt(12 * 4096)


# ------------------------------------------------------------------ autoreset semantics
def test_autoreset_matches_syncvectorenv_semantics(f16, orc):
    """gymnasium 0.29.1 SyncVectorEnv: on done the wrapper resets at once; obs is the RESET
    observation, reward/terminated belong to the terminal step, the terminal observation goes to
    info['final_observation'] under the mask info['_final_info']."""
    n = 1024
    env = f16.F16VecEnv(CFG, n, dtype=torch.float64, autoreset=True, seed=2)
    x_init = env.state.clone()
    acts = torch.ones(n, dtype=torch.float64, device="cuda")      # full aft stick -> wz >= 50 deg/s within ~45 steps
    first_done = None
    prev_state = None
    for k in range(80):
        prev_state = env.state.clone()
        obs, r, term, trunc, info = env.step(acts)
        if term.any():
            first_done = k
            break
    assert first_done is not None and first_done < 70
    m = term.cpu().numpy()
    st = env.state.double().cpu().numpy()
    # terminated envs were re-initialised: counters cleared, state inside the reset box, obs = reset obs
    assert np.all(env.episode_length.cpu().numpy()[m] == 0) and np.all(env.total_return.cpu().numpy()[m] == 0)
    assert np.all(env.episode_length.cpu().numpy()[~m] == first_done + 1)
    assert np.all(np.abs(st[2][m]) <= np.radians(5)) and np.all(st[0][m] == 0) and np.all(st[8][m] == 0)
    o = obs.cpu().numpy()
    ref0 = env.ref_bank_host[env.ref_idx.cpu().numpy(), 0]
    assert np.allclose(o[m, 0], st[1][m] / 35000) and np.allclose(o[m, 3], ref0[m])
    assert np.all(r.cpu().numpy()[m] < -990)                     # terminal-step reward carries the -1000
    # the terminal observation/infos are those of the step that ended the episode (oracle from prev_state)
    idx = np.nonzero(m)[0][:16]
    a1 = np.ones((1, idx.size))
    ref = orc.rollout(prev_state.cpu().numpy().T[idx], a1, env.ref_bank_host, np.zeros(idx.size, dtype=np.int32))
    fo = info["final_observation"].cpu().numpy()[idx]
    assert np.abs(fo[:, :3] - ref["obs"][0][:, :3]).max() < 1e-12
    assert np.all(info["final_info"]["episode_length"].cpu().numpy()[idx] == first_done + 1)
    assert torch.equal(info["_final_info"], term)
    # in-kernel episode statistics (ballot + shuffle reduction): count / mean return / mean length of the finished episodes
    n_fin, mean_ret, mean_len = env.pop_episode_stats()
    assert n_fin == int(m.sum()) and mean_len == first_done + 1
    assert abs(mean_ret - info["final_info"]["total_return"].cpu().numpy()[m].mean()) < 1e-9
    assert env.pop_episode_stats()[0] == 0
    # a second env with the same seed but autoreset off freezes instead
    env2 = f16.F16VecEnv(CFG, n, dtype=torch.float64, autoreset=False, seed=2)
    assert torch.equal(env2.state, x_init)
    for k in range(first_done + 3):
        obs2, r2, term2, _, _ = env2.step(acts)
    assert np.array_equal(term2.cpu().numpy()[m], np.ones(m.sum(), bool))
    assert np.all(r2.cpu().numpy()[m] == 0)


def test_observation_noise_only_touches_wz_channels(f16):
    cfg = dict(CFG, noise=0.1)
    clean = f16.F16VecEnv(CFG, 8192, dtype=torch.float64, seed=4, autoreset=False)
    noisy = f16.F16VecEnv(cfg, 8192, dtype=torch.float64, seed=4, autoreset=False)
    a = torch.zeros(8192, dtype=torch.float64, device="cuda")
    oc, rc, _, _, _ = clean.step(a)
    on, rn, _, _, _ = noisy.step(a)
    assert torch.equal(clean.state, noisy.state) and torch.equal(rc, rn)      # dynamics & reward unaffected
    assert torch.equal(oc[:, [0, 2, 3]], on[:, [0, 2, 3]])
    dwz = ((on[:, 1] - oc[:, 1]) * 2 * np.radians(150)).cpu().numpy()          # back to rad/s
    assert abs(np.degrees(dwz).std() - 0.5) < 0.03 and abs(np.degrees(dwz).mean()) < 0.03
    assert np.allclose((on[:, 4] - oc[:, 4]).cpu().numpy(), -0.5 * dwz, atol=1e-15)


def test_step_host_equals_device_step(f16):
    n = 1 << 17
    a_env = f16.F16VecEnv(CFG, n, seed=9)
    b_env = f16.F16VecEnv(CFG, n, seed=9)
    acts = (torch.rand((3, n)) * 2 - 1)
    for k in range(3):
        o, r, d, _, _ = a_env.step(acts[k].cuda())
        oh, rh, dh = b_env.step_host(acts[k].numpy())
        assert torch.equal(o.cpu(), oh) and torch.equal(r.cpu(), rh[0]) and torch.equal(d.cpu().view(torch.uint8), dh[0])
    assert torch.equal(a_env.state, b_env.state)


@pytest.mark.parametrize("layout", ["aos", "soa"])
def test_step_host_zero_copy_route(f16, layout):
    """Pinned + mapped host buffers: the kernel reads the actions from and bulk-stores obs/reward/done to host
    memory itself.  Ragged n, so the two-envs-per-thread kernel and the one-env remainder kernel both write host
    memory; bit-identical to the device step and to the staged route (pageable actions)."""
    if os.environ.get("F16B200_TWO_PER_THREAD") == "0" or os.environ.get("F16B200_HOST_ZERO_COPY") == "0":
        pytest.skip("debug knob set: the zero-copy route needs the two-envs-per-thread kernel")
    n = (1 << 17) + 77
    a_env = f16.F16VecEnv(CFG, n, seed=9, obs_layout=layout)
    b_env = f16.F16VecEnv(CFG, n, seed=9, obs_layout=layout)
    c_env = f16.F16VecEnv(CFG, n, seed=9, obs_layout=layout)
    acts = (torch.rand((4, n)) * 2 - 1).pin_memory()
    out = (torch.empty((n, 5) if layout == "aos" else (5, n)).pin_memory(), torch.empty((1, n)).pin_memory(),
           torch.empty((1, n), dtype=torch.uint8).pin_memory())
    for k in range(4):
        o, r, d, _, _ = a_env.step(acts[k].cuda())
        oh, rh, dh = b_env.step_host(acts[k], out=out)
        assert b_env.last_host_route == "zero_copy"
        assert torch.equal(o.cpu(), oh) and torch.equal(r.cpu(), rh[0]) and torch.equal(d.cpu().view(torch.uint8), dh[0])
        oc, rc, dc = c_env.step_host(acts[k].numpy().copy())
        assert c_env.last_host_route == "staged"
        assert torch.equal(oc, oh) and torch.equal(rc, rh) and torch.equal(dc, dh)
    assert torch.equal(a_env.state, b_env.state) and torch.equal(a_env.state, c_env.state)


def test_graphed_step_equals_eager_step(f16):
    n = 1 << 16
    a_env = f16.F16VecEnv(CFG, n, seed=9)
    b_env = f16.F16VecEnv(CFG, n, seed=9)
    static = torch.zeros(n, device="cuda")
    graph, (obs_g, rew_g, term_g, _, _) = b_env.make_graphed_step(static)
    for name in ("state", "episode_length", "total_return", "ref_idx", "done", "episode_no"):
        getattr(b_env, name).copy_(getattr(a_env, name))    # undo the warm-up step
    acts = torch.rand((5, n), device="cuda") * 2 - 1
    for k in range(5):
        o, r, d, _, _ = a_env.step(acts[k])
        static.copy_(acts[k])
        graph.replay()
        assert torch.equal(o, obs_g) and torch.equal(r, rew_g) and torch.equal(d, term_g)
    assert torch.equal(a_env.state, b_env.state)


def test_fp32_compensated_positions(f16, golden):
    """With the Kahan carry the fp32 position integrals stay ~1e-7 from the reference; without it
    Ox drifts by up to an ulp per step (the reason the carry exists)."""
    errs = {}
    for comp in (True, False):
        env = _golden_env(f16, golden, torch.float32, compensated_position=comp)
        env.step_k(dev(golden["traj_actions"], torch.float32))
        full = golden["traj_ep_len"] == 1000
        st = env.state.T.cpu().numpy()[full]
        errs[comp] = rel_err(st[:, :2], golden["traj_states"][-1][full][:, :2], 1.0).max()
    print("fp32 position error, compensated / plain:", errs)
    assert errs[True] < 5e-6 and errs[True] <= errs[False]


def test_ragged_batch_sizes(f16, orc):
    """Batch sizes that are not multiples of the warp / CTA size take the scalar store path for the
    tail warp; results must not depend on it."""
    x0, acts = config2_inputs(300, seed=77, T=1000)
    bank = __import__("f16_model_b200").build_reference_bank(0.01, 10, "combo", True)
    ref = orc.rollout(x0, acts[:40], bank, np.zeros(300, dtype=np.int32))
    for n in (1, 31, 33, 257, 300):
        env = f16.F16VecEnv(CFG, n, dtype=torch.float64, autoreset=False, reference_bank=bank)
        env.set_state(x0[:n].T)
        o, r, d = env.step_k(dev(acts[:40, :n]), obs_every_step=True)
        assert np.abs(o.cpu().numpy() - ref["obs"][:, :n]).max() < FP64_RTOL
        assert np.array_equal(d.cpu().numpy(), ref["done"][:, :n])
        assert rel_err(env.state.T.cpu().numpy(), ref["final_state"][:n], FIELD_SCALE).max() < FP64_RTOL


def test_soa_observation_layout(f16):
    a_env = f16.F16VecEnv(CFG, 4096, seed=9, obs_layout="aos")
    b_env = f16.F16VecEnv(CFG, 4096, seed=9, obs_layout="soa")
    acts = torch.zeros(4096, device="cuda")
    oa = a_env.step(acts)[0]
    ob = b_env.step(acts)[0]
    assert ob.shape == (4096, 5) and torch.equal(oa, ob.contiguous())


def test_non_finite_state_terminates_instead_of_trapping(f16):
    """The reference raises ValueError from inside the atmosphere / interpolators on a NaN state
    (SURVEY section 5); the kernel flags the episode as terminated instead."""
    env = f16.F16VecEnv(CFG, 64, dtype=torch.float64, autoreset=False, seed=1)
    st = env.state.clone()
    st[1, 3] = float("nan")       # altitude of env 3
    st[2, 5] = float("nan")       # pitch rate of env 5
    env.set_state(st)
    _, r, d, _, _ = env.step(torch.zeros(64, dtype=torch.float64, device="cuda"))
    d = d.cpu().numpy()
    assert d[3] and d[5] and d.sum() == 2


def test_bad_arguments_raise(f16):
    env = f16.F16VecEnv(CFG, 64)
    with pytest.raises(ValueError):
        env.step(torch.zeros(63, device="cuda"))
    with pytest.raises(TypeError):
        env.step([0.0] * 64)
    with pytest.raises(ValueError):
        f16.F16VecEnv(CFG, 64, dtype=torch.float16)
    with pytest.raises(KeyError):
        f16.F16VecEnv({"dt": 0.01, "tn": 10}, 64)
    with pytest.raises(ValueError):
        env.step_k(torch.zeros(64, device="cuda"))
    with pytest.raises(ValueError):
        f16.F16VecEnv(dict(CFG, init_state=np.zeros((5, 9))), 64)
