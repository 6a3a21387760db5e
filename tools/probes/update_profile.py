"""Probe: kernel-time breakdown of one PPO minibatch step (torch profiler), 65536 envs x 128 steps, 32 minibatches."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
import f16_model_b200 as f16
CFG = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": False, "scenario": "combo"}
envs = f16.F16VecEnv(CFG, 65536, seed=1, fast_math=True)
pol = f16.TensorCorePolicy(seed=1, logstd_init=-1.0)
cfg = f16.PPOConfig(num_steps=128, num_minibatches=32, update_epochs=4, use_cuda_graph=False, fused_update=bool(int(os.environ.get("FUSED", "1"))))
tr = f16.PPOTrainer(envs, pol, cfg)
tr.train(num_updates=1)
adv, ret = tr._adv, tr._ret
mb = torch.randperm(tr.batch_size, device="cuda")[: tr.minibatch_size]
for _ in range(3): tr._minibatch_step(mb)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): tr._minibatch_step(mb)
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print("total device time per minibatch step: %.3f ms" % (tot / 5 / 1e3))
for e in rows[:28]:
    print("%8.1f us x%3d  %5.1f%%  %s" % (e.device_time_total / 5, e.count // 5, 100 * e.device_time_total / tot, e.key[:110]))
