#!/usr/bin/env python
"""Time the closed-loop rollout kernel (f16_rollout) with CUDA events: 65,536 envs x 128 steps and 2^20 envs x 16 steps,
GPU box only."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import f16_model_b200 as f16  # noqa: E402

CFG = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": False, "scenario": "combo"}


def bufs(K, N, dev="cuda"):
    z = lambda *s, **k: torch.zeros(s, device=dev, **k)
    return dict(obs_next=z(K, N, 5), actions=z(K, N), logprobs=z(K, N), values=z(K, N), rewards=z(K, N),
                dones=z(K, N, dtype=torch.uint8), next_value=z(N))


def run(N, K, reps=5):
    env = f16.F16VecEnv(CFG, N, dtype=torch.float32, seed=3, fast_math=True)
    pol = f16.TensorCorePolicy(seed=1, logstd_init=-1.0)
    blob = f16.pack_train_blob16(pol.actor_mean, pol.critic)
    b = bufs(K, N)
    obs0 = env.reset(seed=5)[0].clone()
    ls = pol.actor_logstd.detach()
    f16.fused_rollout(env, blob, ls, obs0, **b, policy_seed=1)
    first = {k: v.clone() for k, v in b.items()}
    ts = []
    for _ in range(reps):
        o = b["obs_next"][K - 1].clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        f16.fused_rollout(env, blob, ls, o, **b, policy_seed=1)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], first


def main():
    out = {}
    for N, K in ((65536, 128), (1 << 20, 16), (4096, 128)):
        us, _ = run(N, K)
        out[f"{N}x{K}"] = {"us": us, "us_per_step": us / K, "env_steps_per_s": N * K / (us * 1e-6)}
    print(json.dumps({k: {a: round(b, 2) if a != 'env_steps_per_s' else float('%.4g' % b) for a, b in v.items()} for k, v in out.items()}))


if __name__ == "__main__":
    main()
