import os, sys, time, torch
sys.path.insert(0, os.getcwd())
import f16_model_b200 as f16
CFG = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": False, "scenario": "combo"}
N = 1 << 20
env = f16.F16VecEnv(CFG, N, seed=1, fast_math=True)
a = (torch.rand((4, N)) * 2 - 1).pin_memory()
out = (torch.empty((N, 5)).pin_memory(), torch.empty((1, N)).pin_memory(), torch.empty((1, N), dtype=torch.uint8).pin_memory())
for s in range(5): env.step_host(a[s % 4], out=out)
torch.cuda.synchronize(); t0 = time.perf_counter()
for s in range(100): env.step_host(a[s % 4], out=out)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(os.environ.get("F16B200_HOST_CHUNKS"), "chunks:", N * 100 / dt / 1e9, "e9 env-steps/s", dt / 100 * 1e3, "ms/step")
