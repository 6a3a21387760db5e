"""The headline kernel (`env_step2_kernel`: fp32, autoreset, two aircraft per thread -- every BENCH / SCALE number)
held to the oracle THROUGH terminations and resets (reference: env/env.py:46-70 step, :103-117 check_state).

Two legs, on BASELINE config-2 inputs (the +-1 stick slams that trip the 50 deg/s limit, the altitude cases, a
ragged batch size so the remainder kernel runs too):

  (a) `env_step2_kernel` == `env_step_kernel<float, ., autoreset, ., .>` BITWISE for 1000 steps, i.e. through every
      termination, stacked penalty, `final_*` store and autoreset, both as K = 1 gym steps and as K-step launches.
      The kernel is selected per env through `f16_config.flags` (F16_FLAG_ONE_ENV_PER_THREAD).
  (b) at every env's FIRST termination the terminal reward (with the stacked -1000 penalties), the
      `final_observation`, `final_info.total_return` and `episode_length` match `orc.rollout` on the same inputs to
      the fp32 tolerance; termination steps agree with the fp64 oracle (an env within rounding of a threshold may
      move by one step).
"""
import numpy as np
import pytest

from workloads import config2_inputs

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

CFG = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": True, "scenario": "combo"}
N = 4096 + 512 + 37          # full 512-env tiles, a full 256-env tile for the K-step kernel's remainder, a ragged tail
T = 1000


@pytest.fixture(scope="module")
def f16():
    import f16_model_b200 as m

    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device; the product has no CPU fallback")
    m.load_library()
    return m


@pytest.fixture(scope="module")
def workload(f16, orc):
    x0, acts = config2_inputs(N, seed=4321)
    # six hand-made envs that cross the 300 m floor AND the 50 deg/s limit on the same step: stacked -2000 penalties
    for j, (wz, a) in enumerate([(0.95, 1.0), (0.95, -1.0), (-0.95, 1.0), (-0.95, -1.0), (0.90, 0.0), (-0.90, 0.0)]):
        i = 16 + j
        x0[i, 1], x0[i, 2], x0[i, 3], x0[i, 4], x0[i, 5] = 300.3, wz, np.radians(-10), 200.0, np.radians(2)
        acts[:, i] = a
    bank = f16.build_reference_bank(0.01, 10, None, False)          # 144 signals
    ref_idx = ((np.arange(N) * 11) % bank.shape[0]).astype(np.int32)
    ref = orc.rollout(x0, acts, bank, ref_idx, want_states=False, want_obs=True)
    return x0, acts, bank, ref_idx, ref


def _make(f16, workload, fast, two_per_thread):
    x0, acts, bank, ref_idx, _ = workload
    env = f16.F16VecEnv(CFG, N, dtype=torch.float32, autoreset=True, seed=77, fast_math=fast, reference_bank=bank,
                        two_per_thread=two_per_thread)
    env.set_state(x0.T, ref_idx=ref_idx)
    return env


def _same(a, b, *names):
    for n in names:
        assert torch.equal(getattr(a, n), getattr(b, n)), n


@pytest.mark.parametrize("fast", [False, True])
def test_gym_step_two_env_kernel_bitwise_and_vs_oracle_through_resets(f16, workload, fast):
    x0, acts, bank, ref_idx, ref = workload
    two, one = _make(f16, workload, fast, True), _make(f16, workload, fast, False)
    assert two._cfg.flags == 0 and one._cfg.flags == 1
    a = torch.tensor(acts, dtype=torch.float32, device="cuda")
    first = np.full(N, -1)                                            # step of each env's first termination
    term_reward = np.zeros(N)
    fin_obs = np.zeros((N, 5))
    fin_ret, fin_len = np.zeros(N), np.zeros(N, dtype=np.int64)
    resets = 0
    for k in range(T):
        o2, r2, d2, _, i2 = two.step(a[k])
        o1, r1, d1, _, i1 = one.step(a[k])
        # (a) bitwise, every step: outputs ...
        assert torch.equal(o2, o1) and torch.equal(r2, r1) and torch.equal(d2, d1), k
        if k % 50 == 49 or bool(d2.any()):                            # ... and the whole persistent state
            _same(two, one, "state", "episode_length", "total_return", "ref_idx", "episode_no")
            assert torch.equal(i2["final_observation"], i1["final_observation"])
            assert torch.equal(i2["final_info"]["total_return"], i1["final_info"]["total_return"])
            assert torch.equal(i2["final_info"]["episode_length"], i1["final_info"]["episode_length"])
        if bool(d2.any()):
            m = d2.cpu().numpy()
            resets += int(m.sum())
            new = m & (first < 0)
            if new.any():
                first[new] = k
                term_reward[new] = r2.cpu().numpy()[new]
                fin_obs[new] = i2["final_observation"].cpu().numpy()[new]
                fin_ret[new] = i2["final_info"]["total_return"].cpu().numpy()[new]
                fin_len[new] = i2["final_info"]["episode_length"].cpu().numpy()[new]
    _same(two, one, "state", "episode_length", "total_return", "ref_idx", "episode_no")
    assert torch.equal(two.episode_stats, one.episode_stats)          # in-kernel statistics (count, sum return, sum length)
    assert first.min() >= 0 and resets > N                           # everybody ended at least once (time-out at the latest)

    # (b) first terminations against the fp64 oracle
    o_first = ref["ep_len"] - 1                                       # oracle: step index of the terminating step
    same_step = first == o_first
    assert same_step.mean() > 0.995, same_step.mean()
    assert np.abs(first - o_first).max() <= 1                         # fp32 rounding may move a threshold crossing by one step
    early = same_step & (o_first < T - 1)
    assert early.sum() > 400, "workload must exercise early terminations"
    idx = np.nonzero(same_step)[0]
    k_ = o_first[idx]
    assert np.array_equal(fin_len[idx], ref["ep_len"][idx])
    o_rew = ref["reward"][k_, idx]
    assert (o_rew[o_first[idx] < T - 1] < -990).all() and (o_rew < -1990).any()   # stacked penalties are exercised (wz + altitude)
    assert np.abs(term_reward[idx] - o_rew).max() <= 1e-4 * 1000      # reward ~ -1000 / -2000: 1e-4 relative
    assert np.abs(fin_obs[idx] - ref["obs"][k_, idx]).max() < 1e-4    # observations are O(1) (normalised)
    tol = 1e-4 * np.maximum(np.abs(ref["total_return"][idx]), 1.0)
    assert (np.abs(fin_ret[idx] - ref["total_return"][idx]) <= tol).all()


@pytest.mark.parametrize("fast", [False, True])
def test_k_step_two_env_kernel_bitwise_and_vs_oracle_through_resets(f16, workload, fast):
    """Same comparison for the K-step instantiation (`env_step2_kernel<., ., SINGLE=false>`), K = 25 per launch."""
    x0, acts, bank, ref_idx, ref = workload
    two, one = _make(f16, workload, fast, True), _make(f16, workload, fast, False)
    a = torch.tensor(acts, dtype=torch.float32, device="cuda")
    K = 25
    first = np.full(N, -1)
    fin_ret, fin_len = np.zeros(N), np.zeros(N, dtype=np.int64)
    term_reward = np.zeros(N)
    for c in range(T // K):
        o2, r2, d2 = two.step_k(a[c * K:(c + 1) * K])
        o1, r1, d1 = one.step_k(a[c * K:(c + 1) * K])
        assert torch.equal(o2, o1) and torch.equal(r2, r1) and torch.equal(d2, d1), c
        _same(two, one, "state", "episode_length", "total_return", "ref_idx", "episode_no", "_final_obs", "_final_len", "_final_ret")
        d = d2.cpu().numpy().astype(bool)
        hit = d.any(axis=0) & (first == -1)
        if hit.any():
            kk = d.argmax(axis=0)
            once = hit & (d.sum(axis=0) == 1)                         # final_* hold the LAST episode that ended in the call
            first[hit] = c * K + kk[hit]
            term_reward[hit] = r2.cpu().numpy()[kk[hit], np.nonzero(hit)[0]]
            fin_ret[once] = two._final_ret.cpu().numpy()[once]
            fin_len[once] = two._final_len.cpu().numpy()[once]
            first[hit & ~once] = -2 - first[hit & ~once]              # seen, but final_* not attributable
    seen = first != -1
    assert seen.all()
    ok = first >= 0
    o_first = ref["ep_len"] - 1
    idx = np.nonzero(ok & (first == o_first))[0]
    assert idx.size > 0.99 * N
    assert np.array_equal(fin_len[idx], ref["ep_len"][idx])
    assert np.abs(term_reward[idx] - ref["reward"][o_first[idx], idx]).max() <= 0.1
    tol = 1e-4 * np.maximum(np.abs(ref["total_return"][idx]), 1.0)
    assert (np.abs(fin_ret[idx] - ref["total_return"][idx]) <= tol).all()
    assert torch.equal(two.episode_stats, one.episode_stats)


def test_flags_are_validated(f16):
    env = f16.F16VecEnv(CFG, 64)
    env._cfg.flags = 1 << 10
    with pytest.raises(ValueError):
        env.step(torch.zeros(64, device="cuda"))


def test_reseed_reaches_a_captured_graph(f16):
    """reset(seed=...) re-keys the RNG through a device scalar, so in-graph autoresets captured earlier draw with the
    NEW seed, exactly like eager steps do."""
    n = 4096
    eager = f16.F16VecEnv(CFG, n, seed=5, fast_math=True)
    graphed = f16.F16VecEnv(CFG, n, seed=5, fast_math=True)
    static = torch.ones(n, device="cuda")                              # full aft stick: everybody terminates within ~70 steps
    graph, (obs_g, rew_g, term_g, _, _) = graphed.make_graphed_step(static)
    for env in (eager, graphed):
        env.reset(seed=123)                                           # re-key AFTER the capture
    assert torch.equal(eager.state, graphed.state)
    n_done = 0
    for k in range(90):
        o, r, d, _, _ = eager.step(static)
        graph.replay()
        assert torch.equal(o, obs_g) and torch.equal(r, rew_g) and torch.equal(d, term_g), k
        n_done += int(d.sum())
    assert n_done >= n
    assert torch.equal(eager.state, graphed.state) and torch.equal(eager.episode_no, graphed.episode_no)
    other = f16.F16VecEnv(CFG, n, seed=5, fast_math=True)
    other.reset(seed=124)
    assert not torch.equal(other.state, eager.init_state)


@pytest.mark.parametrize("two_per_thread", [True, False])
def test_segment_bank_equals_full_row_bank_bitwise(f16, two_per_thread):
    """The default reference bank for random `combo` episodes is the 52 KB segment bank (ref_idx = packed segment rows).
    Against a bank of the 6480 whole rows (26 MB) with the same seed: resets draw the same signal number, so observations
    (which carry wz_ref), rewards, terminations and states are bit-identical, step for step, through resets."""
    cfg = dict(CFG, determenistic_ref=False)
    n = 8192 + 37
    seg = f16.F16VecEnv(cfg, n, seed=31, fast_math=True, two_per_thread=two_per_thread)
    row = f16.F16VecEnv(cfg, n, seed=31, fast_math=True, two_per_thread=two_per_thread, bank_size="full")
    assert seg.ref_mode == "segments" and row.ref_mode == "rows" and seg._cfg.n_ref == row._cfg.n_ref == 6480
    assert tuple(seg.ref_bank.shape) == (300, 48) and tuple(row.ref_bank.shape) == (6480, 1000)
    assert torch.equal(seg.state, row.state)
    assert np.array_equal(seg.reference_signals(), row.reference_signals())
    g = torch.Generator(device="cuda").manual_seed(9)
    acts = torch.rand((1100, n), device="cuda", generator=g) * 2 - 1
    acts[:, ::4] = 1.0                                   # a quarter of the fleet keeps terminating (~45 steps) and drawing new signals
    n_done = 0
    for k in range(1100):                                # past the 1000-step time-out: everybody resets at least once
        os_, rs, ds, _, _ = seg.step(acts[k])
        orow, rr, dr, _, _ = row.step(acts[k])
        assert torch.equal(os_, orow) and torch.equal(rs, rr) and torch.equal(ds, dr), k
        n_done += int(ds.sum()) if k % 100 == 99 else 0
    assert torch.equal(seg.state, row.state) and torch.equal(seg.episode_no, row.episode_no)
    assert int(seg.episode_no.min()) >= 2 and n_done > 0
    assert np.array_equal(seg.reference_signals(), row.reference_signals())
    # K-step launches read the segment bank the same way
    o1, r1, d1 = seg.step_k(acts[:64])
    o2, r2, d2 = row.step_k(acts[:64])
    assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(d1, d2)
