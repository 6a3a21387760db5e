import numpy as np, torch, sys, os, json, time
sys.path.insert(0, os.getcwd())
import f16_model_b200 as f16
Vg, Hg = np.meshgrid(np.linspace(100, 300, 41), np.linspace(1000, 10000, 37))
starts = f16.reference_starts()
f16.trim_grid(Vg[:2], Hg[:2], starts=starts[::16], device="cuda")
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter()
    r = f16.trim_grid(Vg, Hg, starts=starts, device="cuda", return_all=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"cells": int(Vg.size), "starts": int(starts.shape[0]), "seconds": dt, "searches_per_s": Vg.size * starts.shape[0] / dt,
                      "iterations": float(np.sum(r["all_iterations"])), "accepted": float(np.mean(r["cost"] <= 1e-5)),
                      "checksum": float(np.sum(r["stab"]) + np.sum(r["throttle"]) + np.sum(r["alpha"]))}))
