#!/usr/bin/env python
"""One-off, larger validation of the C oracle against the UNMODIFIED reference (build container only:
needs /root/reference; uses the stand-ins in oracle/shims).  Flies N random episodes (config-2 style
inputs) through the reference's F16 env and through the oracle, reports the worst deviations.
    python tests/studies/validate_oracle_vs_reference.py [n_envs=256]
The committed goldens (tests/golden) hold 12 such episodes; this run is the wider check quoted in DESIGN.md."""
import multiprocessing as mp
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "oracle", "shims"), os.environ.get("F16_REFERENCE", "/root/reference"), ROOT,
                os.path.join(ROOT, "tests")]
warnings.simplefilter("ignore")


def fly(job):
    from F16model.env import F16

    i, x0, acts, scenario = job
    cfg = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": True, "scenario": scenario,
           "init_state": np.asarray(x0[:6], dtype=float), "init_control": np.array([x0[6], 0.0])}
    env = F16(cfg)
    env.reset()
    S = np.zeros((1000, 9)); R = np.zeros(1000); k_done = 1000
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        for k in range(1000):
            _, r, d, _, info = env.step(np.array([acts[k]]))
            S[k], R[k] = info["state"].to_array(), r
            if d:
                k_done = k + 1
                break
    return i, S, R, k_done, env.ref_signal.wz_ref


def main():
    from workloads import config2_inputs
    from oracle import f16_oracle as orc

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    x0, acts = config2_inputs(n, seed=4321)
    scen = ["combo", "cos", "step", "pure_step"]
    jobs = [(i, x0[i], acts[:, i], scen[i % 4]) for i in range(n)]
    t0 = time.time()
    with mp.Pool(min(8, os.cpu_count())) as pool:
        res = sorted(pool.map(fly, jobs, chunksize=4), key=lambda r: r[0])
    refs = np.stack([r[4] for r in res])
    out = orc.rollout(x0, acts, refs, np.arange(n), want_obs=False)
    worst_abs = worst_rel = 0.0
    steps = 0
    scale = np.array([1.0, 1.0, 1e-2, 1e-2, 1.0, 1e-2, 1e-2, 1e-2, 1e-2])
    mism = 0
    for i, S, R, k_done, _ in res:
        steps += k_done
        mism += int(out["ep_len"][i] != k_done)
        d = np.abs(out["states"][:k_done, i] - S[:k_done])
        worst_abs = max(worst_abs, d.max())
        worst_rel = max(worst_rel, (d / np.maximum(np.abs(S[:k_done]), scale)).max())
    print(f"{n} episodes, {steps} env-steps of the unmodified reference in {time.time() - t0:.0f} s: "
          f"step-count mismatches {mism}, early terminations {sum(r[3] < 1000 for r in res)}, "
          f"worst |d| {worst_abs:.3e}, worst relative {worst_rel:.3e}")


if __name__ == "__main__":
    main()
