"""The fp32 drift bound is stated as 1e-4 after 1000 steps.  One env of the config-2 workload sits above it in the fp32 kernels
(1.2e-4).  This CPU test shows that the excess is a property of the MODEL, not of the kernels: the fp64 oracle itself, in exact fp64
arithmetic, with nothing but its nine state values rounded to fp32 between steps -- the minimum any fp32 implementation does --
leaves the pure-fp64 trajectory by about as much on exactly that env, while the typical env moves by 1e-6.  The GPU test
(test_gpu_parity.py::test_config2_fp32_drift_bound) uses the same per-env figure to decide which envs the 1e-4 bound binds."""
import numpy as np

from workloads import config2_inputs, fp32_storage_sensitivity


def test_fp32_state_storage_alone_breaks_1e4_on_the_sensitive_env(orc):
    import f16_model_b200 as f16

    x0, acts = config2_inputs(4096)
    bank = f16.build_reference_bank(0.01, 10, None, False)
    ref_idx = (np.arange(4096) * 7) % bank.shape[0]
    ref = orc.rollout(x0, acts, bank, ref_idx, want_states=False, want_obs=False)
    runs = fp32_storage_sensitivity(orc, x0, acts, bank, ref_idx, ref)          # [3, n]: storage only; storage + half-ulp jitter (two seeds)
    live = np.isfinite(runs[0])
    assert live.sum() > 3400                                      # the envs that fly the whole episode
    storage = runs[0][live]
    assert np.median(storage) < 3e-6 and np.quantile(storage, 0.999) < 3e-5      # typical env: fp32 state storage costs ~1e-6 ...
    assert int(np.nanargmax(runs[0])) == 1226                     # ... the env the fp32 kernels are worst on (tests/studies/fp32_error_study.py)
    assert runs[0][1226] > 8e-5                                   # is at 88 % of the bound from storage rounding alone, in exact fp64 arithmetic,
    assert np.nanmax(runs[1:, 1226]) > 1e-4                       # and past it once each step carries a half-ulp of right-hand-side rounding
    worst = np.nanmax(runs, axis=0)[live]
    assert 1 <= int((worst >= 5e-5).sum()) <= 16                  # a handful of rounding-sensitive envs in 3544
