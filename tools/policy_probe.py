"""Policy forward timing (CUDA graph of 20 back-to-back launches, CUDA events): policy_forward_kernel (TF32) vs
policy_forward16_kernel (fp16 operands, the rollout kernel's schedule).  python tools/policy_probe.py [sizes...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import f16_model_b200 as f16

sizes = [int(a) for a in sys.argv[1:]] or [4096, 16384, 65536, 1 << 18, 1 << 20]
pol = f16.TensorCorePolicy(seed=1, logstd_init=-1.0)
for n in sizes:
    obs = torch.rand((n, 5), device="cuda") * 2 - 1
    out = tuple(torch.empty(n, device="cuda") for _ in range(3))
    rec = {"n": n}
    for prec in ("tf32", "fp16"):
        for k in range(3):
            pol.act(obs, step=k, out=out, precision=prec)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for k in range(20):
                pol.act(obs, step=k, out=out, precision=prec)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 200
        rec[prec + "_us"] = round(us, 2)
        rec[prec + "_tanh_per_clk_per_sm"] = round(256.0 * n / (us * 1e-6) / 148 / 1.965e9, 2)
    print(json.dumps(rec), flush=True)
