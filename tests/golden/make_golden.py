#!/usr/bin/env python
"""Generate tests/golden/ref_golden.npz by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference).  The reference's four
missing third-party imports are satisfied by the stand-ins in oracle/shims/
(ambiance is a restatement of ICAO-1993 -- parity is UNPINNED at that boundary,
see oracle/shims/ambiance/__init__.py; the other three contribute no
arithmetic).  Everything else -- numpy, scipy, the reference's own code and
tables -- is the real thing.

    python tests/golden/make_golden.py            # ~3 min on 8 cores

The fixture pins (a) the C oracle (tests/test_oracle_golden.py, CPU) and (b)
the CUDA path directly (tests/test_gpu_parity.py, GPU box, where
/root/reference does not exist).
"""
import multiprocessing as mp
import os
import random
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("F16_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(ROOT, "oracle", "shims"), REF]
warnings.simplefilter("ignore")

from F16model.data import atmosphere, coeff, plane, thrust  # noqa: E402
from F16model.env import F16, ReferenceSignal, get_trimmed_state_control  # noqa: E402
from F16model.env.env import get_random_state  # noqa: E402
from F16model.model import Control, States  # noqa: E402
from F16model.model import engine  # noqa: E402
from F16model.model.ODE_3DoF import solve  # noqa: E402

DT, TN, T = 0.01, 10, 1000


# ----------------------------------------------------------------------------
def golden_pointwise(rng):
    g = {}
    # coefficient points: random in/out of range + exact grid nodes + edges
    import F16model.data._aerodynamics as A

    al = np.concatenate([rng.uniform(-0.7, 1.9, 260), A.alpha1, [A.alpha1[0] - 0.2, A.alpha1[-1] + 0.2],
                         np.radians([2.797, 7.3, 48.0, -22.0, 59.999, 60.0, 60.001, 75.0])])
    n = al.size
    fi = rng.uniform(-0.6, 0.6, n)
    fi[260:265] = A.fi1
    fi[265:272] = A.fi3
    fi[-8:] = np.radians([-4.3674, 12.5, 26.0, -27.0, 14.999, 15.0, 20.0, 25.0])
    wz = rng.uniform(-1.1, 1.1, n)
    V = rng.uniform(60.0, 400.0, n)
    out = np.empty((n, 3))
    for i in range(n):
        a = (al[i], 0, fi[i], plane.lef, wz[i], V[i], plane.b_a, plane.sb)
        out[i] = [float(coeff.get_Cx(*a)), float(coeff.get_Cy(*a)), float(coeff.get_Mz(*a))]
    g["coef_in"] = np.stack([al, fi, wz, V], 1)
    g["coef_out"] = out

    import F16model.data._engine as E

    H = np.concatenate([rng.uniform(-200, 31000, 200), E.H1, [3000.0, 0.0, 20000.0, 15240.0]])
    n = H.size
    M = rng.uniform(0.0, 1.6, n)
    M[200:206] = E.M1
    Pa = rng.uniform(0.0, 130.0, n)
    Pa[200:203] = E.Pa1
    M[-4:] = [0.6087, 0.0, 1.2, 1.0]
    Pa[-4:] = [24.462898, 100.0, 75.0, 100.0]
    g["thrust_in"] = np.stack([H, M, Pa], 1)
    g["thrust_out"] = np.array([float(thrust.get_thrust(h, m, p)) for h, m, p in zip(H, M, Pa)])

    h = np.concatenate([np.linspace(-1000, 10900, 60), np.linspace(11100, 31000, 30), [0.0, 3000.0, 11000.0, 11019.0, 11020.0]])
    g["isa_in"] = h
    g["isa_out"] = np.array([[float(atmosphere.get_density(v)), float(atmosphere.get_speed_of_sound(v))] for v in h])

    pa = rng.uniform(0, 110, 200)
    th = rng.uniform(-0.2, 1.2, 200)
    pa[:8] = [49.999, 50.0, 50.001, 0.0, 100.0, 25.0, 75.0, 35.0]
    th[:8] = [0.77, 0.7699, 0.7701, 0.0, 1.0, 0.77, 0.2, 1.0]
    g["power_in"] = np.stack([pa, th], 1)
    g["power_out"] = np.array([float(engine.power_level(p, t)) for p, t in zip(pa, th)])
    thr = np.concatenate([np.linspace(0, 1, 21), [0.3767, 0.2884, 0.77]])
    g["thrpos_in"] = thr
    g["thrpos_out"] = np.array([engine.find_correct_thrust_position(t) for t in thr])

    a = np.concatenate([rng.uniform(-1.3, 1.3, 100), [-1.0, 1.0, 0.0, -0.17469281, 1e-9]])
    g["rescale_in"] = a
    g["rescale_out"] = np.array([float(F16.rescale_action(np.array([v]))) for v in a])

    # full RHS
    n = 300
    x = np.stack([
        rng.uniform(0, 5000, n), rng.uniform(400, 10500, n), rng.uniform(-0.9, 0.9, n), rng.uniform(-0.8, 1.2, n),
        rng.uniform(80, 320, n), rng.uniform(-0.45, 1.3, n), rng.uniform(-0.5, 0.5, n), rng.uniform(-1.3, 1.3, n),
        rng.uniform(0, 100, n)], 1)
    x[:40, 1] = rng.uniform(11500, 29500, 40)   # upper ISA layers (unpinned restatement)
    u = np.stack([rng.uniform(-0.5, 0.5, n), rng.uniform(-0.1, 1.1, n)], 1)
    x[0] = [0, 3456.7, 0.05, 0.03, 180, 0.06, -0.02, 0.1, 30]
    u[0] = [0.1, 0.5]
    dx = np.empty_like(x)
    for i in range(n):
        dx[i] = solve(States(*x[i]), Control(*u[i])).to_array()
    g["rhs_x"], g["rhs_u"], g["rhs_dx"] = x, u, dx
    return g


# ----------------------------------------------------------------------------
def golden_refsignals():
    g = {}
    for sc in ["step", "cos", "pure_step", "combo"]:
        g[f"ref_det_{sc}"] = ReferenceSignal(DT, TN, determenistic=True, scenario=sc).theta_ref
        for seed in (1, 2, 3, 11, 12345):
            random.seed(seed)
            g[f"ref_rnd_{sc}_{seed}"] = ReferenceSignal(DT, TN, determenistic=False, scenario=sc).theta_ref
    for seed in (4, 5, 6, 7, 8, 9):
        random.seed(seed)
        g[f"ref_rnd_None_{seed}"] = ReferenceSignal(DT, TN, determenistic=False, scenario=None).theta_ref
    # other horizons / steps used by the reference's scripts (test/env_test.py: dt=0.01, tn=5;
    # combo needs tn >= 10)
    g["ref_det_pure_step_tn5"] = ReferenceSignal(0.01, 5, determenistic=True, scenario="pure_step").theta_ref
    g["ref_det_step_tn20_dt02"] = ReferenceSignal(0.02, 20, determenistic=True, scenario="step").theta_ref
    g["ref_det_combo_tn20"] = ReferenceSignal(0.01, 20, determenistic=True, scenario="combo").theta_ref
    # random initial states as drawn by reset(seed) (env/env.py:155-158,221-243)
    rs = []
    for seed in (1, 2, 3, 100):
        np.random.seed(seed)
        rs.append(get_random_state().to_array())
    g["random_state_seeds"] = np.array([1, 2, 3, 100])
    g["random_state_out"] = np.array(rs)
    x0, u0 = get_trimmed_state_control()
    g["trim_x0"], g["trim_u0"] = x0, u0
    # airframe constants and bounds (data/plane.py)
    names = ["m", "l", "S", "b_a", "Jz", "rcgx", "Tstab", "Xistab", "maxabsstab", "maxabsdstab", "minthrottle", "maxthrottle", "lef", "sb"]
    g["plane_scalars"] = np.array([float(getattr(plane, k)) for k in names])
    g["plane_state_bound"] = np.array([plane.state_bound[k] for k in ("Oy", "wz", "V")], dtype=float)
    g["plane_random_state_bound"] = np.array([plane.random_state_bound[k] for k in ("Oy", "wz", "V", "alpha", "maxabsstab")], dtype=float)
    return g


# ----------------------------------------------------------------------------
def action_program(kind, rng, u0):
    """Normalised stick sequences [T] (float64), reproducible from (kind, rng)."""
    trim = u0[0] / np.radians(25)
    if kind == "doublet":   # BASELINE config 1: +-2 deg elevator doublet about trim
        a = np.full(T, trim)
        a[100:200] = (u0[0] + np.radians(2)) / np.radians(25)
        a[200:300] = (u0[0] - np.radians(2)) / np.radians(25)
        return a
    if kind == "hold25":    # piecewise constant, held 25 steps
        return np.repeat(rng.uniform(-0.3, 0.1, T // 25), 25)
    if kind == "iid":
        return rng.uniform(-1.0, 1.0, T)
    if kind == "slam+":
        a = np.repeat(rng.uniform(-0.3, 0.1, T // 25), 25)
        a[int(rng.integers(50, 400)):] = 1.0
        return a
    if kind == "slam-":
        a = np.repeat(rng.uniform(-0.3, 0.1, T // 25), 25)
        a[int(rng.integers(50, 400)):] = -1.0
        return a
    if kind == "overrange":  # outside [-1,1]: exercises the clip in rescale_action
        return rng.uniform(-1.6, 1.6, T) * (rng.uniform(0, 1, T) < 0.1) + trim
    raise ValueError(kind)


def run_episode(job):
    """One reference episode from an explicit init state; frozen after done."""
    idx, x0, u_thr, scenario, acts = job
    cfg = {"dt": DT, "tn": TN, "debug_state": False, "determenistic_ref": True, "scenario": scenario,
           "init_state": np.asarray(x0[:6], dtype=float), "init_control": np.array([x0[6], u_thr])}
    env = F16(cfg)
    obs0, _ = env.reset()
    init = env.init_state.to_array()
    S = np.zeros((T, 9)); O = np.zeros((T, 5)); R = np.zeros(T); D = np.zeros(T, dtype=np.uint8)
    k_done = T
    for k in range(T):
        o, r, d, _, info = env.step(np.array([acts[k]]))
        S[k], O[k], R[k], D[k] = info["state"].to_array(), o, r, d
        if d:
            k_done = k + 1
            break
    S[k_done:] = S[k_done - 1]; O[k_done:] = O[k_done - 1]; D[k_done:] = 1
    return idx, init, obs0, env.ref_signal.wz_ref, S, O, R, D, env.episode_length, env.total_return


def golden_trajectories(rng):
    x0t, u0 = get_trimmed_state_control()
    jobs = []
    # 0: BASELINE config 1 (trim, det. pure_step, doublet)
    jobs.append((x0t, u0[1], "pure_step", action_program("doublet", rng, u0)))
    kinds = ["hold25", "hold25", "hold25", "iid", "slam+", "slam-", "overrange", "hold25", "iid"]
    scen = ["combo", "cos", "step", "combo", "combo", "pure_step", "cos", "combo", "step"]
    for kind, sc in zip(kinds, scen):
        al = rng.uniform(*plane.random_state_bound["alpha"])
        x0 = np.array([0.0, rng.uniform(*plane.random_state_bound["Oy"]), rng.uniform(*plane.random_state_bound["wz"]),
                       al, rng.uniform(*plane.random_state_bound["V"]), al, 0.0, 0.0, 0.0])
        jobs.append((x0, 0.0, sc, action_program(kind, rng, u0)))
    # altitude terminations: start just above the floor in a dive / just below the ceiling in a climb
    x_low = np.array([0.0, 305.0, 0.0, np.radians(-8.0), 180.0, np.radians(2.0), 0.0, 0.0, 0.0])
    jobs.append((x_low, 0.0, "combo", action_program("hold25", rng, u0)))
    x_high = np.array([0.0, 29990.0, 0.0, np.radians(20.0), 250.0, np.radians(3.0), 0.0, 0.0, 0.0])
    jobs.append((x_high, 0.9, "combo", action_program("hold25", rng, u0)))
    jobs = [(i,) + j for i, j in enumerate(jobs)]
    with mp.Pool(min(8, len(jobs))) as pool:
        res = sorted(pool.map(run_episode, jobs), key=lambda r: r[0])
    g = {
        "traj_actions": np.stack([j[4] for j in jobs], 1),                # [T, n]
        "traj_x0": np.stack([r[1] for r in res]),                        # [n, 9]
        "traj_obs0": np.stack([r[2] for r in res]),                      # [n, 5]
        "traj_ref": np.stack([r[3] for r in res]),                       # [n, T] wz_ref
        "traj_states": np.stack([r[4] for r in res], 1),                 # [T, n, 9]
        "traj_obs": np.stack([r[5] for r in res], 1),                    # [T, n, 5]
        "traj_reward": np.stack([r[6] for r in res], 1),                 # [T, n]
        "traj_done": np.stack([r[7] for r in res], 1),                   # [T, n]
        "traj_ep_len": np.array([r[8] for r in res], dtype=np.int32),
        "traj_return": np.array([r[9] for r in res]),
    }
    return g


def main():
    rng = np.random.default_rng(20261019)
    g = {}
    g.update(golden_pointwise(rng))
    g.update(golden_refsignals())
    g.update(golden_trajectories(rng))
    import scipy

    g["meta"] = np.array([f"numpy {np.__version__}; scipy {scipy.__version__}; ambiance restated (unpinned); "
                          f"reference /root/reference unmodified; dt={DT} tn={TN}"])
    out = os.path.join(ROOT, "tests", "golden", "ref_golden.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out) // 1024, "KiB;", "ep_len", g["traj_ep_len"], "returns", g["traj_return"])


if __name__ == "__main__":
    main()
