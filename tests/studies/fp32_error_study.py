#!/usr/bin/env python
"""Where does the fp32 mode's error against the fp64 reference come from?  (GPU box)

Runs the BASELINE config-2 inputs (4096 envs x 1000 steps) in fp32 (accurate and fast-math)
and compares every 50th state with the CPU oracle; prints the error distribution, the worst
envs with their stick program, and the error growth over time.  Evidence for the stated fp32
bound in DESIGN.md."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import f16_model_b200 as f16  # noqa: E402
from oracle import f16_oracle as orc  # noqa: E402
from workloads import FP32_FIELD_SCALE, config2_inputs, rel_err  # noqa: E402

CFG = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": True, "scenario": "combo"}
NAMES = ["Ox", "Oy", "wz", "theta", "V", "alpha", "stab", "dstab", "Pa"]


def main():
    n = 4096
    x0, acts = config2_inputs(n)
    bank = f16.build_reference_bank(0.01, 10, None, False)
    ref_idx = ((np.arange(n) * 7) % bank.shape[0]).astype(np.int32)
    ref = orc.rollout(x0, acts, bank, ref_idx, want_states=True, want_obs=False)
    full = ref["ep_len"] == 1000
    # classify stick programs: std of per-step action differences
    dact = np.abs(np.diff(acts, axis=0)).mean(axis=0)
    amax = np.abs(acts).max(axis=0)
    for fast in (False, True):
        for comp in (True, False):
            env = f16.F16VecEnv(CFG, n, dtype=torch.float32, autoreset=False, reference_bank=bank, fast_math=fast,
                                compensated_position=comp)
            env.set_state(x0.T, ref_idx=ref_idx)
            a = torch.tensor(acts, dtype=torch.float32, device="cuda")
            errs = []
            for c in range(20):
                env.step_k(a[c * 50:(c + 1) * 50])
                st = env.state.T.double().cpu().numpy()
                errs.append(rel_err(st, ref["states"][c * 50 + 49], FP32_FIELD_SCALE))
            errs = np.stack(errs)            # [20, n, 9]
            ok = full & (env.episode_length.cpu().numpy() == 1000)
            e_end = errs[-1][ok]
            print(f"\n=== fast_math={fast} compensated_position={comp}: {ok.sum()} full-episode envs")
            print("max per field  :", " ".join(f"{nm}={v:.2e}" for nm, v in zip(NAMES, e_end.max(axis=0))))
            print("99.9% per field:", " ".join(f"{nm}={v:.2e}" for nm, v in zip(NAMES, np.quantile(e_end, 0.999, axis=0))))
            print("median         :", " ".join(f"{nm}={v:.2e}" for nm, v in zip(NAMES, np.median(e_end, axis=0))))
            print("envs over 1e-4 (any field):", int((e_end.max(axis=1) > 1e-4).sum()), " over 5e-5:", int((e_end.max(axis=1) > 5e-5).sum()))
            worst = np.argsort(-errs[-1].max(axis=1) * ok)[:6]
            for j in worst:
                f = int(errs[-1][j].argmax())
                grow = " ".join(f"{errs[t][j].max():.1e}" for t in (0, 3, 7, 11, 15, 19))
                s = ref["states"][:, j]
                print(f"  env {j}: worst field {NAMES[f]} {errs[-1][j].max():.2e}; |a|max {amax[j]:.2f} mean|da| {dact[j]:.3f}; "
                      f"V {x0[j,4]:.0f}; alpha range [{np.degrees(s[:,5].min()):.1f},{np.degrees(s[:,5].max()):.1f}] deg; "
                      f"wz max {np.degrees(np.abs(s[:,2]).max()):.1f} deg/s; growth@50,200,400,600,800,1000: {grow}")


if __name__ == "__main__":
    main()
