#!/usr/bin/env python
"""Time the fused PPO gradient kernel (f16_ppo_grad16 + reduction) on config 4's minibatch (262,144 rows gathered from a
65,536 x 128 rollout) with a CUDA graph of 10 launches, and the whole update (128 minibatches).  GPU box only."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if len(sys.argv) > 1 and sys.argv[1].startswith("lib"):
    os.environ["F16B200_LIB"] = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "f16-model_b200", sys.argv[1], "libf16b200.so")
import f16_model_b200 as f16  # noqa: E402
from f16_model_b200 import _capi  # noqa: E402

CFG = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": False, "scenario": "combo"}


def graph_us(fn, reps=10, replays=5):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(reps):
            fn()
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * replays)


def main():
    n4, T4 = 65536, 128
    envs = f16.F16VecEnv(CFG, n4, dtype=torch.float32, seed=1, fast_math=True)
    pol = f16.TensorCorePolicy(seed=1, logstd_init=-1.0)
    tr = f16.PPOTrainer(envs, pol, f16.PPOConfig(num_steps=T4, num_minibatches=32, update_epochs=4))
    hist = tr.train(num_updates=8)
    L = _capi.load_library()
    st = lambda: _capi.stream_ptr(torch.device("cuda"))
    b = tr._flat
    mb = tr._perm_buf[: tr.minibatch_size]
    packed = "unpacked" not in sys.argv
    srcs = (tr._rows.data_ptr(), None, None, None, tr._rv.data_ptr(), None) if packed else (
        b["obs"].data_ptr(), b["actions"].data_ptr(), b["logprobs"].data_ptr(), b["adv"].data_ptr(), b["ret"].data_ptr(), b["val"].data_ptr())
    args16 = (mb.numel(), tr._train_blob.data_ptr(), *srcs, mb.data_ptr(), pol.actor_logstd.data_ptr(),
              tr._adv_stats.data_ptr(), tr._g.data_ptr(), 0.5, 0.5, 0.0, 1, 1)
    us = graph_us(lambda: _capi.check(L.f16_ppo_grad16(*args16, tr._grad_scratch.data_ptr(), st())))
    print(json.dumps({"grad16_us_per_minibatch": us, "packed_rows": packed, "rows": mb.numel(), "update_s": hist[-1]["update_s"], "update_s_all": [round(h["update_s"], 5) for h in hist], "rollout_s": hist[-1]["rollout_s"],
                      "value_loss": hist[-1]["value_loss"], "approx_kl": hist[-1]["approx_kl"]}))


    if hasattr(L, "f16_debug_grad_trace"):
        import ctypes as C
        import numpy as np
        out = np.zeros((8, 4, 10), dtype=np.uint64)
        L.f16_debug_grad_trace.argtypes = [C.c_void_p]
        assert L.f16_debug_grad_trace(out.ctypes.data) == 0
        t = out.astype(np.int64)
        if hasattr(L, "f16_debug_epoch_trace"):
            et = np.zeros((4, 12), dtype=np.uint64)
            L.f16_debug_epoch_trace.argtypes = [C.c_void_p]
            assert L.f16_debug_epoch_trace(et.ctypes.data) == 0
            names = ["tile loop", "sync", "flush+prefetch", "meet 1", "reduce+norm share", "meet 2", "Adam+scatter", "meet 3", "reload"]
            for r, who in enumerate(("CTA 0 slot 0", "CTA 0 slot 1", "last CTA slot 0", "last CTA slot 1")):
                d = np.diff(et[r, :10].astype(np.int64))
                print(f"epoch kernel, minibatch 2, {who} (clk): " + "  ".join(f"{n}={v}" for n, v in zip(names, d)))
        if hasattr(L, "f16_debug_epoch_cta"):
            ct = np.zeros((160, 12), dtype=np.uint64)
            L.f16_debug_epoch_cta.argtypes = [C.c_void_p]
            assert L.f16_debug_epoch_cta(ct.ctypes.data) == 0
            ct = ct[:148, :10].astype(np.int64)
            ct -= ct[:, 0].min()
            print("per-CTA globaltimer (ns) relative to the earliest start of minibatch 2's tile loop; columns = min / median / max over 148 CTAs")
            for i, nm in enumerate(["loop start", "loop end", "synced", "flushed (arrive 1)", "met 1", "reduced (arrive 2)", "met 2", "stepped (arrive 3)", "met 3", "reloaded"]):
                print(f"  {nm:22s} {ct[:, i].min():7d} {int(np.median(ct[:, i])):7d} {ct[:, i].max():7d}   argmax CTA {int(ct[:, i].argmax())}")
            dur = ct[:, 1] - ct[:, 0]
            print("  tile-loop duration per CTA (ns): min / median / max", dur.min(), int(np.median(dur)), dur.max(), " slowest CTAs:", np.argsort(dur)[-8:].tolist())
            fl = ct[:, 3] - ct[:, 2]
            print("  flush duration per CTA (ns): min / median / max", fl.min(), int(np.median(fl)), fl.max())
        if int(os.environ.get("F16B200_GRAD16", "0")) >= 0:   # ppo_grad16w_kernel: rows 4 s + (warp & 3), points 0..8 of rounds 2..5
            names = ["retire+S1", "wait L1", "S2", "wait L2", "S3", "token+issue S3", "wait D3", "S4+token+issue"]
            d = np.diff(t[:, :, :9], axis=2)
            print("period of a round (clk), slot 0 / slot 1:", np.diff(t[0, :, 0]).mean().round(), np.diff(t[4, :, 0]).mean().round())
            print("slot 1 behind slot 0 at the top of a round (clk):", (t[4, :, 0] - t[0, :, 0]).tolist())
            for w in range(8):
                print(f"slot {w // 4} warp {w % 4}: " + "  ".join(f"{n}={d[w, :, i].mean():.0f}" for i, n in enumerate(names)))
            return
        d = np.diff(t[:, :, :8], axis=2)
        names = ["retire+S1 x2+gather", "S2(0)", "S2(1)", "S3(0)", "S3(1)", "S4(0)", "S4(1)"]
        print("period of a PAIR of tiles (clk):", np.diff(t[:, :, 0], axis=1).mean(axis=1).round())
        for w in range(8):
            print(f"warp {w}: " + "  ".join(f"{n}={d[w, :, i].mean():.0f}" for i, n in enumerate(names)))


if __name__ == "__main__":
    main()
