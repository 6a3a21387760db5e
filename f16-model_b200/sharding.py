"""Multi-GPU sharding of the env batch: envs are independent, so each rank
(one process per GPU) owns a contiguous block and the simulation needs NO collective.
``torch.distributed`` is used only by the callers that aggregate statistics."""


def shard_range(total_envs, rank, world_size):
    """Contiguous block [begin, end) of global env ids owned by ``rank``; blocks differ by at
    most one env and cover [0, total_envs) exactly."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, rem = divmod(int(total_envs), int(world_size))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def make_sharded_env(config, total_envs, rank, world_size, device, **kw):
    """This rank's ``F16VecEnv`` over its block of a ``total_envs`` global batch; RNG streams
    are keyed by global env id, so the union over ranks equals the single-GPU batch."""
    from .env import F16VecEnv

    begin, end = shard_range(total_envs, rank, world_size)
    return F16VecEnv(config, end - begin, device=device, env_id_base=begin, **kw)


def global_episode_stats(env, group=None):
    """All-reduced (sum of returns, sum of lengths, episodes finished this step) over ranks --
    the only cross-rank traffic of a rollout, at logging cadence."""
    import torch
    import torch.distributed as dist

    if hasattr(env, "episode_stats"):          # kernel-side accumulators (sum of returns, sum of lengths, count)
        s = env.episode_stats
        stats = torch.stack([s[1], s[2], s[0]]).double()
    else:
        done = env._done_out[0].bool()
        stats = torch.stack([
            (env._final_ret * done).sum().double(),
            (env._final_len * done).sum().double(),
            done.sum().double(),
        ])
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats
