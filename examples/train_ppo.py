#!/usr/bin/env python
"""PPO training on the batched F-16 env -- the counterpart of the reference's
``python control/ppo/ppo_train_gsde.py`` (flags as in control/ppo/utils.py:11-142), with the whole
rollout on the device.  Single GPU:  python examples/train_ppo.py --num-envs 65536
Multi GPU:   torchrun --nproc-per-node 8 examples/train_ppo.py --num-envs 65536   (envs per GPU)"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import f16_model_b200 as f16  # noqa: E402
from f16_model_b200.sharding import make_sharded_env, shard_range  # noqa: E402

ENV_CONFIG = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": False, "scenario": "combo"}


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--exp-name", default="F16")
    p.add_argument("--learning-rate", type=float, default=3e-4)
    p.add_argument("--seed", type=int, default=1)
    p.add_argument("--total-timesteps", type=int, default=50_000_000)
    p.add_argument("--num-envs", type=int, default=65536)
    p.add_argument("--num-steps", type=int, default=128)
    p.add_argument("--anneal-lr", type=int, default=1)
    p.add_argument("--gamma", type=float, default=0.99)
    p.add_argument("--gae-lambda", type=float, default=0.95)
    p.add_argument("--num-minibatches", type=int, default=32)
    p.add_argument("--update-epochs", type=int, default=4)
    p.add_argument("--norm-adv", type=int, default=1)
    p.add_argument("--clip-coef", type=float, default=0.2)
    p.add_argument("--clip-vloss", type=int, default=1)
    p.add_argument("--ent-coef", type=float, default=0.0)
    p.add_argument("--vf-coef", type=float, default=0.5)
    p.add_argument("--max-grad-norm", type=float, default=0.5)
    p.add_argument("--target-kl", type=float, default=None)
    p.add_argument("--save-dir", default="runs/")
    p.add_argument("--tensorboard", type=int, default=0)
    return p.parse_args()


def main():
    args = parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    run_name = f"{args.exp_name}__{args.seed}__{int(time.time())}"
    envs = make_sharded_env(ENV_CONFIG, args.num_envs * world, rank, world, dev, seed=args.seed, fast_math=True)
    agent = f16.TensorCorePolicy(device=dev, seed=args.seed, env_id_base=shard_range(args.num_envs * world, rank, world)[0])
    cfg = f16.PPOConfig(learning_rate=args.learning_rate, seed=args.seed, total_timesteps=args.total_timesteps // world,
                        num_steps=args.num_steps, anneal_lr=bool(args.anneal_lr), gamma=args.gamma, gae_lambda=args.gae_lambda,
                        num_minibatches=args.num_minibatches, update_epochs=args.update_epochs, norm_adv=bool(args.norm_adv),
                        clip_coef=args.clip_coef, clip_vloss=bool(args.clip_vloss), ent_coef=args.ent_coef, vf_coef=args.vf_coef,
                        max_grad_norm=args.max_grad_norm, target_kl=args.target_kl)
    trainer = f16.PPOTrainer(envs, agent, cfg)
    tb = f16.artifacts.TensorboardLogger(args.save_dir, run_name, vars(args)) if (args.tensorboard and rank == 0) else None

    def log(s):
        if rank == 0:
            print(f"update {s['update']:4d}  step {s['global_step'] * world:>12d}  reward/step {s['mean_step_reward']:.4f}  "
                  f"episodes {s['episodes']}  v_loss {s['value_loss']:.3f}  kl {s['approx_kl']:.4f}  "
                  f"rollout {s['rollout_sps'] * world:.3g} steps/s  update {s['update_s']:.2f} s", flush=True)
            if tb:
                tb(s)

    trainer.train(log=log)
    if rank == 0:
        print("saved:", f16.artifacts.save_agent(agent, args.save_dir, run_name))
    trainer.close()                 # releases the captured graphs and the peer-memory exchange buffers (must precede destroy_process_group)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
