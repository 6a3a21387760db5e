"""Trim search (BASELINE config 5b).  CPU: the oracle's Nelder-Mead restatement is
pinned iterate-for-iterate against scipy.optimize.fmin, and reproduces the two trim points the
reference publishes.  GPU: the batched device search against the oracle on the same (cell, start)
pairs, the published trims, and a whole altitude/speed grid."""
import numpy as np
import pytest
import scipy.optimize

PUBLISHED = [  # (V, H, stab deg, throttle, alpha deg): env/env.py:248-257 ; control/ppo/pretty_plots.ipynb cell 13
    (200.0, 3000.0, -4.3674, 0.3767, 2.7970),
    (175.0, 2500.0, -4.4712, 0.2884, 3.4334),
]
STARTS = np.array([[np.radians(-4), 0.45, np.radians(2)], [np.radians(10), 0.7, np.radians(-6)], [0.0, 0.2, 0.0],
                   [np.radians(-20), 1.45, np.radians(10)], [np.radians(5), 0.95, np.radians(-2)]])


def test_oracle_nelder_mead_is_scipy_fmin(orc):
    """Same cost function, same starts -> identical iterates: final x bit-equal, same iteration count."""
    for V, H in [(200.0, 3000.0), (120.0, 6000.0)]:
        f = lambda u: orc.trim_cost(V, H, np.asarray(u, dtype=float))          # noqa: E731
        x, cost, it = orc.trim([V], [H], STARTS[:3])
        for s in range(3):
            out = scipy.optimize.fmin(f, STARTS[s].copy(), xtol=1e-10, ftol=1e-10, maxfun=5e3, maxiter=1e3, disp=False,
                                      full_output=True)
            assert np.array_equal(out[0], x[0, s]) and out[1] == cost[0, s] and out[2] == it[0, s]


def test_oracle_reproduces_published_trims(orc):
    for V, H, stab, thr, alpha in PUBLISHED:
        x, cost, _ = orc.trim([V], [H], STARTS)
        b = int(np.argmin(cost[0]))
        assert cost[0, b] <= 1e-5                                             # the reference's acceptance
        assert abs(np.degrees(x[0, b, 0]) - stab) < 5e-5 and abs(x[0, b, 1] - thr) < 5e-5
        assert abs(np.degrees(x[0, b, 2]) - alpha) < 5e-5


@pytest.mark.gpu
def test_device_trim_matches_oracle_per_start(orc):
    import f16_model_b200 as f16

    V = np.array([200.0, 175.0, 120.0, 250.0])
    H = np.array([3000.0, 2500.0, 6000.0, 9000.0])
    starts = f16.reference_starts()[::29]                                      # 94 of the 2706 reference starts
    r = f16.trim_grid(V, H, starts=starts, return_all=True)
    x, cost, it = orc.trim(V, H, starts)
    ok = np.isfinite(cost)
    same_path = (r["all_iterations"] == it) & ok
    assert same_path.mean() > 0.9                                              # libm differs in the last ulp: a few searches branch differently
    assert np.max(np.abs(r["all_x"][same_path] - x[same_path])) < 1e-7
    assert np.max(np.abs(r["all_cost"][same_path] - cost[same_path])) < 1e-9
    # whatever the path, both pick the same trim per cell (first accepted start in the given order)
    for c in range(V.size):
        b = int(np.argmax(cost[c] <= 1e-5))
        assert cost[c, b] <= 1e-5 and r["accepted"][c]
        assert np.max(np.abs(np.array([r["stab"][c], r["throttle"][c], r["alpha"][c]]) - x[c, b])) < 1e-6


@pytest.mark.gpu
def test_device_trim_published_points_and_grid():
    import f16_model_b200 as f16

    r = f16.trim_grid([p[0] for p in PUBLISHED], [p[1] for p in PUBLISHED])    # the reference's full 2706-start grid
    for k, (V, H, stab, thr, alpha) in enumerate(PUBLISHED):
        assert r["accepted"][k] and r["n_starts"] == 2706
        assert abs(np.degrees(r["stab"][k]) - stab) < 5e-5 and abs(r["throttle"][k] - thr) < 5e-5
        assert abs(np.degrees(r["alpha"][k]) - alpha) < 5e-5
    ctrl, alpha, theta = f16.find_trim_position.run(200, 3000, max_starts=64)
    assert abs(np.degrees(ctrl.stab) + 4.3674) < 5e-5 and alpha == theta
    # altitude / speed grid (<= 11 km geopotential, where the ISA restatement uses exact constants)
    Vg, Hg = np.meshgrid(np.linspace(100, 300, 9), np.linspace(1000, 10000, 10))
    g = f16.trim_grid(Vg.ravel(), Hg.ravel(), max_starts=128)
    assert g["accepted"].mean() > 0.9
    acc = g["accepted"]
    assert np.all(np.abs(g["alpha"][acc]) < np.radians(30)) and np.all((g["throttle"][acc] > 0) & (g["throttle"][acc] < 1.5))
    # slower / higher needs more angle of attack
    a = g["alpha"].reshape(Hg.shape)
    assert np.all(a[:, 0] > a[:, -1])
