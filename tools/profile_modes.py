#!/usr/bin/env python
"""Launch the env-step kernel in its two other regimes so that ncu can capture them (GPU box):
   (1) fp32, K = 64 substeps per launch, 2^20 envs   -- issue-bound throughput mode
   (2) fp64, K = 16 substeps per launch, 2^18 envs   -- parity mode
Also prints CUDA-event throughput of both.  Usage under ncu:
   ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s <skip> -c 2 -o out python tools/profile_modes.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import f16_model_b200 as f16  # noqa: E402

CFG = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": False, "scenario": "combo"}


def run(dtype, n, K, reps, fast):
    env = f16.F16VecEnv(CFG, n, dtype=dtype, seed=1, fast_math=fast, bank_size=64)
    acts = torch.rand((K, n), device="cuda", dtype=dtype) * 2 - 1
    out = (env._obs, torch.empty((K, n), device="cuda", dtype=dtype), torch.empty((K, n), dtype=torch.uint8, device="cuda"))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        env.step_k(acts, out=out)
    tot = 0.0
    for r in range(reps):
        flush.fill_(r)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        env.step_k(acts, out=out)
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return n * K * reps / (tot * 1e-3), tot / reps


def main():
    v32, ms32 = run(torch.float32, 1 << 20, 64, 3, True)
    v64, ms64 = run(torch.float64, 1 << 18, 16, 3, False)
    print(json.dumps({"fp32_K64_env_steps_per_s": v32, "fp32_K64_ms_per_launch": ms32,
                      "fp64_K16_env_steps_per_s": v64, "fp64_K16_ms_per_launch": ms64}))


if __name__ == "__main__":
    main()
