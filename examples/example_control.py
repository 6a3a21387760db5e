#!/usr/bin/env python
"""Open-loop elevator step from trim -- the counterpart of the reference's control/example_control.py:
single F-16 env (same NumPy-in / NumPy-out signature as F16model.env.F16), running on the GPU in fp64."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from f16_model_b200 import F16, get_trimmed_state_control  # noqa: E402

x0, u0 = get_trimmed_state_control()
CONFIG = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": True, "scenario": "pure_step",
          "init_state": x0, "init_control": u0}

env = F16(CONFIG)
obs, _ = env.reset()
total = 0.0
for k in range(int(env.tn / env.dt)):
    stab = u0[0] + (np.radians(1.0) if k * env.dt >= 1.0 else 0.0)           # 1 deg elevator step at t = 1 s
    obs, reward, done, _, info = env.step(np.array([stab / np.radians(25)]))
    total += reward
    if k % 100 == 99 or done:
        s = info["state"]
        print(f"t={info['clock']:5.2f}s  H={s.Oy:8.2f} m  wz={np.degrees(s.wz):7.3f} deg/s  theta={np.degrees(s.theta):7.3f} deg  "
              f"alpha={np.degrees(s.alpha):6.3f} deg  reward={reward:.4f}")
    if done:
        break
print("episode length", info["episode_length"], "total return", round(total, 4))
