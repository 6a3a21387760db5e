"""Small end-to-end exercise of every step-kernel route for compute-sanitizer (memcheck / racecheck / synccheck):
one-env and two-env kernels, gym step and K-step, ragged tail, zero-copy host route with bulk-stored observations,
policy forward, PPO rollout + fused update (fp16 and tf32 gradient kernels, Adam, GAE), trim.  Sizes are tiny because the tools slow execution 10-100x.
(compute-sanitizer is closed on this GPU pool -- the gpurun wrapper refuses it -- so here the script only serves as a
plain run-through of every route; the bounds evidence is the parity suite itself: ragged sizes 1..300, canary-free
bit-exact comparisons against the oracle, and sharding invariance at 8 Mi envs.)"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
import f16_model_b200 as f16
CFG = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": False, "scenario": "combo"}
n = 4 * 296 * 256 // 8 + 77            # several tiles per CTA on a few CTAs would need a big n; keep one wave + ragged tail
env = f16.F16VecEnv(CFG, n, seed=1, fast_math=True)
a = torch.rand((8, n), device="cuda") * 2 - 1
for k in range(3):
    env.step(a[k])
env.step_k(a[:4])
env.step_k(a[4:], obs_every_step=True)
ah = (torch.rand(n) * 2 - 1).pin_memory()
out = (torch.empty((n, 5)).pin_memory(), torch.empty((1, n)).pin_memory(), torch.empty((1, n), dtype=torch.uint8).pin_memory())
env.step_host(ah, out=out)
assert env.last_host_route == "zero_copy"
env.step_host(ah.numpy().copy())
e64 = f16.F16VecEnv(CFG, 1000, seed=2, dtype=torch.float64, autoreset=False)
e64.step(torch.zeros(1000, dtype=torch.float64, device="cuda"))
pol = f16.TensorCorePolicy(seed=1)
pol.act(env._obs, env=env)
f16.trim_grid([200.0], [3000.0])
CFG_D = dict(CFG, determenistic_ref=True)
for prec in ("fp16", "tf32"):
    envs = f16.F16VecEnv(CFG_D, 512, seed=3)
    tr = f16.PPOTrainer(envs, f16.TensorCorePolicy(seed=3, logstd_init=-1.0),
                        f16.PPOConfig(num_steps=8, num_minibatches=2, update_epochs=1, fused_precision=prec, use_cuda_graph=False))
    tr.train(num_updates=1)
torch.cuda.synchronize()
print("sanitize workload done", float(env.total_return.sum()))
