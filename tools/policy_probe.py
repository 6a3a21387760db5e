import os, sys, torch
sys.path.insert(0, os.getcwd())
import f16_model_b200 as f16
pol = f16.TensorCorePolicy(seed=1)
for n in (128 * 148, 128 * 296, 65536, 1 << 20):
    obs = torch.rand((n, 5), device="cuda") * 2 - 1
    for _ in range(3): pol.act(obs, step=0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(20): pol.act(obs, step=k)
    e1.record(); torch.cuda.synchronize()
    print(n, "envs:", e0.elapsed_time(e1) / 20 * 1e3, "us per launch", n / (e0.elapsed_time(e1) / 20 * 1e-3) / 1e9, "e9 envs/s")
