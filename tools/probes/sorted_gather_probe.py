#!/usr/bin/env python
"""Does the order of the rows INSIDE a minibatch matter to the gradient kernel's gathers?  Times the persistent epoch kernel
(f16_ppo_epoch16, 32 minibatch steps of 262,144 rows gathered from a 65,536 x 128 rollout) with the epoch's permutation as drawn
(random order inside each minibatch) and with every minibatch's index segment sorted (same partition into minibatches, rows visited
in increasing address order).  GPU box only."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import f16_model_b200 as f16  # noqa: E402

CFG = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": False, "scenario": "combo"}


def time_us(fn, reps=3, replays=4):
    for _ in range(2):
        fn()
    gr = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
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
    envs = f16.F16VecEnv(CFG, 65536, dtype=torch.float32, seed=1, fast_math=True)
    pol = f16.TensorCorePolicy(seed=1, logstd_init=-1.0)
    tr = f16.PPOTrainer(envs, pol, f16.PPOConfig(num_steps=128, num_minibatches=32, update_epochs=4))
    tr.train(num_updates=3)
    n_mb, mbs = tr.batch_size // tr.minibatch_size, tr.minibatch_size
    step = lambda: tr._fused_minibatch_step(tr._perm_buf[:mbs], adv_stats=tr._adv_stats_epoch[0], epoch_minibatches=n_mb)  # noqa: E731
    out = {}
    out["random_order_us_per_epoch"] = time_us(step)
    seg = tr._perm_buf.view(n_mb, mbs)
    seg.copy_(torch.sort(seg, dim=1).values)
    out["sorted_inside_minibatch_us_per_epoch"] = time_us(step)
    out["us_per_minibatch"] = {k: v / n_mb for k, v in out.items()}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
