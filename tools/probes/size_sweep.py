"""Probe: gym-step time against batch size (one resident batch per size, CUDA graph of 20 steps): separates the
per-launch overhead (ramp-up / drain / dependency gap) from the asymptotic per-env cost."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
import f16_model_b200 as f16
CFG = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": False, "scenario": "combo"}
for lg in (16, 17, 18, 19, 20, 21, 22, 23):
    N = 1 << lg
    env = f16.F16VecEnv(CFG, N, seed=1, fast_math=True)
    pool = torch.rand((20, N), device="cuda") * 2 - 1      # a fresh stick per step, as in bench.py
    for k in range(3): env.step(pool[k])
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g):
            for k in range(20): env.step(pool[k])
    torch.cuda.current_stream().wait_stream(side)
    reps = max(3, (1 << 26) // N // 20)
    for _ in range(2): g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * 20)
    print(f"n=2^{lg} {us:8.2f} us/step  {N / us * 1e-3:7.2f} e9 env-steps/s  {113 * N / us * 1e-3 / 6532.2:5.3f} of roofline", flush=True)
    del env, g
