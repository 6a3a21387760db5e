"""Host-buffer path probe: env-steps/s of F16VecEnv.step_host for the chunk schedule in the environment
(F16B200_HOST_CHUNKS, F16B200_HOST_RATIO), next to the raw pinned device->host copy rate of the same bytes."""
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
print("chunks", os.environ.get("F16B200_HOST_CHUNKS"), "ratio", os.environ.get("F16B200_HOST_RATIO"), ":",
      round(N * 100 / dt / 1e9, 4), "e9 env-steps/s", round(dt / 100 * 1e3, 4), "ms/step")
if os.environ.get("F16B200_RAW_COPY"):
    dev = torch.empty(N * 25, dtype=torch.uint8, device="cuda")
    host = torch.empty(N * 25, dtype=torch.uint8).pin_memory()
    for _ in range(3): host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50): host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("raw pinned D2H of 25 B x 2^20:", round(N * 25 * 50 / dt / 1e9, 2), "GB/s", round(dt / 50 * 1e3, 4), "ms per copy")
