"""One trim_grid call on the bench's grid (21 x 19 cells x 677 starts = 270 k searches) -- the target of the ncu capture of trim_kernel."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import f16_model_b200 as f16
Vg, Hg = np.meshgrid(np.linspace(100, 300, 21), np.linspace(1000, 10000, 19))
starts = f16.reference_starts()[::4]
r = f16.trim_grid(Vg, Hg, starts=starts, device="cuda", return_all=True)
torch.cuda.synchronize()
print(json.dumps({"searches": int(Vg.size * starts.shape[0]), "iterations": float(np.sum(r["all_iterations"]))}))
