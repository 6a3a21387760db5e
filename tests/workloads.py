"""Seeded synthetic workloads shared by the GPU parity tests (BASELINE config 2)."""
import numpy as np

T = 1000


def config2_inputs(n=4096, seed=1234, T=T):
    """Random initial states per plane.random_state_bound, plus stick programs:
    3/4 piecewise-constant U(-0.3,0.1) held 25 steps, 1/8 i.i.d. U(-1,1), 1/8 sustained +-1
    from a random onset (forces the 50 deg/s termination), and a few hand-made states next to
    the altitude limits.  Everything fp64, generated on the host."""
    rng = np.random.default_rng(seed)
    d5 = np.radians(5)
    al = rng.uniform(-d5, d5, n)
    x0 = np.zeros((n, 9))
    x0[:, 1] = rng.uniform(2000, 4000, n)
    x0[:, 2] = rng.uniform(-d5, d5, n)
    x0[:, 3] = al
    x0[:, 4] = rng.uniform(90, 250, n)
    x0[:, 5] = al
    acts = np.repeat(rng.uniform(-0.3, 0.1, (T // 25, n)), 25, axis=0)
    kind = rng.permutation(n) % 8
    iid = kind == 6
    acts[:, iid] = rng.uniform(-1, 1, (T, int(iid.sum())))
    slam = np.nonzero(kind == 7)[0]
    onset = rng.integers(20, 800, slam.size)
    sign = rng.choice([-1.0, 1.0], slam.size)
    for j, o, s in zip(slam, onset, sign):
        acts[o:, j] = s
    # altitude terminations
    k = min(8, n // 4)
    x0[:k, 1] = np.linspace(301.0, 340.0, k)
    x0[:k, 3] = np.radians(-10.0)
    x0[:k, 5] = np.radians(2.0)
    x0[k:2 * k, 1] = np.linspace(29950.0, 29999.0, k)
    x0[k:2 * k, 3] = np.radians(25.0)
    x0[k:2 * k, 4] = 250.0
    x0[k:2 * k, 5] = np.radians(3.0)
    return x0, acts


FIELD_SCALE = np.array([1.0, 1.0, 1e-2, 1e-2, 1.0, 1e-2, 1e-2, 1e-2, 1e-2])
"""Per-field floor of the relative error |d| / max(|ref|, scale): many fields pass through 0
(wz, dstab ~ 1e-16 at trim), so 'relative' needs a floor (SURVEY section 7, hard parts)."""


FP32_FIELD_SCALE = np.array([1.0, 1.0, 0.05, 0.05, 1.0, 0.05, 0.05, 0.1, 1.0])
"""Floors of the STATED fp32 bound (1e-4 after 1000 steps): about 5 % of each field's working
range -- 0.05 rad/s (~3 deg/s) for wz, 0.05 rad for the angles, 0.1 rad/s for the actuator rate.
fp32 state stored in HBM is rounded once per env-step (a few ulp of drift per step in the
integrated angles), which the fp64 mode does not have; DESIGN.md lists the measured maxima
under both floor sets."""


def rel_err(a, ref, scale):
    a, ref = np.asarray(a, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return np.abs(a - ref) / np.maximum(np.abs(ref), scale)


def fp32_storage_sensitivity(orc, x0, acts, bank, ref_idx, ref, T=1000):
    """How far does the fp64 MODEL move when nothing but its state is kept in fp32?  The oracle is stepped one step at a time in its
    exact fp64 arithmetic; between steps the nine state values are rounded to fp32 (run 0), and additionally jittered by a uniform
    +-2^-24 relative error (runs 1, 2: the half-ulp a rounded fp32 right-hand side would add).  Returns [3, n]: per run and env the
    deviation of the final state from the pure-fp64 trajectory in the units of the fp32 bound (FP32_FIELD_SCALE floors); NaN for
    envs that do not fly the whole episode in every run.  (tests/studies/fp32_error_study.py
    has the per-field breakdown; this is the assertion-grade version.)"""
    f32 = lambda a: np.asarray(a, dtype=np.float64).astype(np.float32).astype(np.float64)
    a32 = f32(acts)
    out = np.zeros((3, x0.shape[0]))
    ok = ref["ep_len"] == T
    for run in range(3):
        rng = np.random.default_rng(run) if run else None
        x = f32(x0)
        alive = np.ones(x0.shape[0], dtype=bool)
        for k in range(T):
            r = orc.rollout(x, a32[k:k + 1], bank, ref_idx, want_states=False, want_obs=False)   # the state does not depend on the reference row
            alive &= r["ep_len"] == 1
            x = f32(r["final_state"])
            if rng is not None:
                x = x * (1.0 + rng.uniform(-1.0, 1.0, x.shape) * 2.0 ** -24)
        ok &= alive
        out[run] = rel_err(x, ref["final_state"], FP32_FIELD_SCALE).max(axis=1)
    return np.where(ok[None], out, np.nan)
