#!/usr/bin/env python
"""Timings for the BASELINE configs that are not the headline bench line (GPU box):
  config 4: PPO rollout + update with 65536 envs on device (control/ppo MLP policy)
  config 5b: batched trim-state search over an altitude/speed grid
  plus a short "does PPO learn" run.  Prints one JSON object per section.
    python tools/config_bench.py [--updates 3] [--learn-updates 40]
Under torchrun the envs are sharded over ranks and the PPO gradients all-reduced over NCCL."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import f16_model_b200 as f16  # noqa: E402
from f16_model_b200.sharding import make_sharded_env, shard_range, trim_grid_sharded  # noqa: E402

CFG = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": False, "scenario": "combo"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--updates", type=int, default=3)
    ap.add_argument("--learn-updates", type=int, default=0)
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--num-steps", type=int, default=128)
    ap.add_argument("--tf32-update", type=int, default=0)
    ap.add_argument("--peer", type=int, default=1, help="0: NCCL gradient all-reduce instead of the peer-memory one fused into the reduction kernel")
    ap.add_argument("--only-ppo", action="store_true")
    ap.add_argument("--learn-fused", type=int, default=1, help="0: torch autograd update in the learning run (comparison curve)")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)

    # ---- config 4 -------------------------------------------------------------------------
    envs = make_sharded_env(CFG, args.envs * world, rank, world, dev, seed=1, fast_math=True)
    pol = f16.TensorCorePolicy(device=dev, seed=1, logstd_init=-1.0, env_id_base=shard_range(args.envs * world, rank, world)[0])
    cfg = f16.PPOConfig(num_steps=args.num_steps, num_minibatches=32, update_epochs=4, tf32_update=bool(args.tf32_update), peer_allreduce=bool(args.peer))
    tr = f16.PPOTrainer(envs, pol, cfg)
    hist = tr.train(num_updates=args.updates)
    # the gradient all-reduce on its own: 128 eager NCCL calls on the 37 KB flat gradient, as the update issues them
    allreduce_us = None
    if world > 1 and getattr(tr, "_g", None) is not None:
        g = tr._g.clone()
        for _ in range(10):
            torch.distributed.all_reduce(g)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(128):
            torch.distributed.all_reduce(g)
        e1.record()
        torch.cuda.synchronize()
        allreduce_us = e0.elapsed_time(e1) * 1e3 / 128
    # parameters must be identical on every rank after the updates (same gradients, same optimiser state)
    flat = torch.cat([p.detach().reshape(-1) for p in pol.parameters()])
    same = True
    if world > 1:
        ref = flat.clone()
        torch.distributed.broadcast(ref, 0)
        ok = torch.tensor([1 if torch.equal(ref, flat) else 0], device=dev)
        torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN)
        same = bool(ok.item())
    epoch_kernel = bool(getattr(tr, "_step16", None) is not None and cfg.fused_epoch and tr._update_graph and (world == 1 or tr._peer is not None))
    if rank == 0:
        h = hist[-1]
        print(json.dumps({"config": "4: PPO rollout+update, %d envs/GPU x %d steps, %d GPU(s)" % (args.envs, args.num_steps, world),
                          "rollout_env_steps_per_s": h["rollout_sps"] * world, "rollout_s": h["rollout_s"], "update_s": h["update_s"],
                          "minibatches_per_update": cfg.num_minibatches * cfg.update_epochs, "value_loss": h["value_loss"],
                          "approx_kl": h["approx_kl"],
                          "grad_allreduce": ((("peer memory (NVLink), inside the reduction phase of the persistent epoch kernel (f16_ppo_epoch16)" if epoch_kernel else
                                               "peer memory (NVLink), inside ppo_reduce_adam_kernel (reduction + all-reduce + clip + Adam + blob scatter in one launch)"
                                               if getattr(tr, "_step16", None) is not None else "peer memory (NVLink), fused into ppo_grad_reduce_peer_kernel")
                                              if tr._peer is not None else "nccl")
                                             + (", captured in the update graph" if tr._update_graph else ", eager")) if world > 1 else None,
                          "update_s_all": [round(x["update_s"], 6) for x in hist],
                          "optimizer": ("inside the persistent epoch kernel (f16_ppo_epoch16: one cooperative launch per epoch)" if epoch_kernel else
                                        "fused into the reduction kernel (f16_ppo_step16)" if getattr(tr, "_step16", None) is not None else "ppo_adam_kernel (one CTA)"),
                          "nccl_allreduce_us_per_call_37KB": allreduce_us,
                          "allreduce_share_of_update": (allreduce_us * cfg.num_minibatches * cfg.update_epochs * 1e-6 / h["update_s"]) if allreduce_us else None,
                          "parameters_identical_on_all_ranks": same}), flush=True)

    tr.close()                                  # the update graph holds NCCL work: release it before the process group goes away
    if args.only_ppo:
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    # policy forward alone (tensor-core kernel): CUDA events around replays of a graph of 20 launches (an eager Python
    # loop is bound by the interpreter + ctypes call, ~20 us per launch, not by the kernel)
    obs = torch.rand((args.envs, 5), device=dev) * 2 - 1
    for _ in range(3):
        pol.act(obs, step=0)
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph):
            for k in range(20):
                keep = pol.act(obs, step=k)
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(2):
        graph.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 200
    flops = 2 * (8 * 128 + 2 * 64 * 64) * args.envs          # tensor-core MACs actually issued (K padded to 8)
    if rank == 0:
        print(json.dumps({"kernel": "policy_forward_kernel (tcgen05 tf32), graph of 20 launches", "envs": args.envs, "us_per_launch": ms * 1e3,
                          "envs_per_s": args.envs / (ms * 1e-3), "tensor_tflops": flops / (ms * 1e-3) / 1e12}), flush=True)

    # ---- config 5b --------------------------------------------------------------------------
    Vg, Hg = np.meshgrid(np.linspace(100, 300, 41), np.linspace(1000, 10000, 37))
    b, e = shard_range(Vg.size, rank, world)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    g = trim_grid_sharded(Vg, Hg, rank, world, device=dev)     # cells striped over ranks + one all_gather
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    res = torch.tensor(np.stack([g["stab"], g["throttle"], g["alpha"], g["cost"]], 1))
    if rank == 0:
        print(json.dumps({"config": "5b: trim grid %d cells x %d starts (reference grid), %d GPU(s)" % (Vg.size, g["n_starts"], world),
                          "seconds_rank0": dt, "searches_per_s": (e - b) * g["n_starts"] * world / dt,
                          "accepted_fraction": float((res[:, 3] <= 1e-5).double().mean())}), flush=True)

    # ---- learning curve ---------------------------------------------------------------------
    if args.learn_updates and world == 1:
        det = dict(CFG, determenistic_ref=True)
        envs = f16.F16VecEnv(det, 4096, device=dev, seed=3, fast_math=True)
        pol = f16.TensorCorePolicy(device=dev, seed=3, logstd_init=-1.0)
        tr = f16.PPOTrainer(envs, pol, f16.PPOConfig(num_steps=250, num_minibatches=16, update_epochs=5, learning_rate=1e-3,
                                                      clip_coef=0.2, anneal_lr=False, fused_update=bool(args.learn_fused)))
        hist = tr.train(num_updates=args.learn_updates)
        curve = [round(h["mean_step_reward"], 4) for h in hist]
        print(json.dumps({"ppo_learning": "4096 envs x 250 steps per update, det. combo reference", "update": "fused kernels (f16_ppo_epoch16)" if args.learn_fused else "torch autograd",
                          "update_s_last": hist[-1]["update_s"], "mean_step_reward_per_update": curve,
                          "explained_variance_last": hist[-1]["explained_variance"]}), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
