"""Policy-forward probe: time per launch for several batch sizes, eager and replayed from a CUDA graph of 20 launches
(the graph figure excludes the Python / ctypes call overhead)."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
import f16_model_b200 as f16
pol = f16.TensorCorePolicy(seed=1)
for n in (128 * 148, 128 * 296, 65536, 1 << 18, 1 << 20):
    obs = torch.rand((n, 5), device="cuda") * 2 - 1
    out = None
    for _ in range(3): pol.act(obs, step=0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for k in range(20): pol.act(obs, step=k)
    e1.record(); torch.cuda.synchronize()
    eager = e0.elapsed_time(e1) / 20 * 1e3
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g):
            for k in range(20): keep = pol.act(obs, step=k)
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(2): g.replay()
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 200 * 1e3
    flops = n * 2 * 2 * (8 * 64 + 64 * 64)          # tensor-core layers only, both nets
    print(f"{n:8d} envs: eager {eager:7.2f} us, graph {us:7.2f} us per launch = {n / us * 1e-3:6.2f} e9 envs/s, {flops / us * 1e-6:6.1f} TFLOP/s (tensor layers)", flush=True)
