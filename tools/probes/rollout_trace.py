#!/usr/bin/env python
"""Phase timeline of the rollout kernel (debug build: make -C f16-model_b200/csrc EXTRA=-DF16_ROLLOUT_TRACE OUT=../lib_trace).
Prints, for CTA 0, each warp's phase durations (clocks) over steps 8..15 of one launch."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ["F16B200_LIB"] = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "f16-model_b200", sys.argv[1] if len(sys.argv) > 1 else "lib_trace", "libf16b200.so")
import f16_model_b200 as f16  # noqa: E402
from f16_model_b200 import _capi  # noqa: E402

CFG = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": False, "scenario": "combo"}
N, K = 65536, 128
env = f16.F16VecEnv(CFG, N, dtype=torch.float32, seed=3, fast_math=True)
pol = f16.TensorCorePolicy(seed=1, logstd_init=-1.0)
blob = f16.pack_train_blob16(pol.actor_mean, pol.critic)
z = lambda *s, **k: torch.zeros(s, device="cuda", **k)
b = dict(obs_next=z(K, N, 5), actions=z(K, N), logprobs=z(K, N), values=z(K, N), rewards=z(K, N), dones=z(K, N, dtype=torch.uint8), next_value=z(N))
obs0 = env.reset(seed=5)[0].clone()
for _ in range(2):
    f16.fused_rollout(env, blob, pol.actor_logstd.detach(), obs0, **b, policy_seed=1)
torch.cuda.synchronize()
L = _capi.load_library()
out = np.zeros((16, 8, 12), dtype=np.uint64)
L.f16_debug_rollout_trace.argtypes = [C.c_void_p]
rc = L.f16_debug_rollout_trace(out.ctypes.data)
assert rc == 0, rc
t = out.astype(np.int64)
t0 = t[:, 0, 0].min()
names = ["pack+sync", "wait MMA1", "E1 actor", "E1 critic+syncs", "wait MMA2a", "E2", "head+env"]
print("step period per warp (clk):", np.diff(t[:, :, 0], axis=1).mean(axis=1).round())
iss = t[:, :, 8:]
t = t[:, :, :8]
d = np.diff(t, axis=2)                 # [warp][step][7]
print("mean phase durations over steps 8..15 (clk):")
for w in range(16):
    print(f"warp {w:2d} (group {w // 4}, smsp {w % 4}): " + "  ".join(f"{n}={d[w, :, i].mean():6.0f}" for i, n in enumerate(names)))
for w in range(16):
    if iss[w].any():
        print(f"issuer warp {w}: 4 MMAs issue={np.diff(iss[w], axis=1)[:, 0].mean():.0f}  bias MMA={np.diff(iss[w], axis=1)[:, 1].mean():.0f}  commit={np.diff(iss[w], axis=1)[:, 2].mean():.0f} clk")
print("timeline of step 10, SMSP 0 warps (start of each phase relative to the earliest):")
for w in (0, 4, 8, 12):
    print(f"warp {w:2d}: " + " ".join(f"{int(x - t0):7d}" for x in t[w, 2, :]) + " | next " + " ".join(f"{int(x - t0):7d}" for x in t[w, 3, :3]))
print("start of every traced step, SMSP 0 warps, modulo the mean period:")
P = np.diff(t[:, :, 0], axis=1).mean()
for w in (0, 4, 8, 12):
    print(f"warp {w:2d}: " + " ".join(f"{int((x - t0) % P):6d}" for x in t[w, :, 0]))

st = np.zeros((16, 128), dtype=np.uint64)
L.f16_debug_rollout_starts.argtypes = [C.c_void_p]
assert L.f16_debug_rollout_starts(st.ctypes.data) == 0
st = st.astype(np.int64)
print("start of step k relative to warp 0's start of the same step (clk), SMSP 0 warps 4, 8, 12; and warp 0's step period:")
for k in list(range(0, 32, 2)) + list(range(32, 128, 8)):
    print(f"k={k:3d}: " + " ".join(f"{int(st[w, k] - st[0, k]):7d}" for w in (4, 8, 12)) + f"   period {int(st[0, min(k + 1, 127)] - st[0, min(k + 1, 127) - 1]):6d}")
print("total clocks warp 0, steps 0..127:", int(st[0, 127] - st[0, 0]))
