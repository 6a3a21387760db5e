"""World-size-2 test of the multi-process host logic on CPU (gloo): each rank owns the
contiguous env block `shard_range` gives it, nothing is exchanged on the data path, and the
only collective -- the episode-statistics all-reduce at logging cadence -- reproduces the
single-process numbers."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from f16_model_b200.sharding import global_episode_stats, shard_range

TOTAL = 1001      # deliberately not divisible by the world size


class _FakeEnvShard:
    """The three output tensors `global_episode_stats` reads, for one rank's block."""

    def __init__(self, begin, end):
        rng = np.random.default_rng(7)
        done = rng.random(TOTAL) < 0.3
        ret = rng.normal(size=TOTAL) * 100
        length = rng.integers(1, 1000, TOTAL)
        self._done_out = torch.tensor(done[begin:end], dtype=torch.uint8)[None]
        self._final_ret = torch.tensor(ret[begin:end])
        self._final_len = torch.tensor(length[begin:end], dtype=torch.int32)
        self.expected = torch.tensor([ret[done].sum(), float(length[done].sum()), float(done.sum())], dtype=torch.float64)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    begin, end = shard_range(TOTAL, rank, world)
    shard = _FakeEnvShard(begin, end)
    stats = global_episode_stats(shard)
    ok = torch.allclose(stats, shard.expected, rtol=1e-12)
    spans = [None] * world
    dist.all_gather_object(spans, (begin, end))
    cover = spans[0][0] == 0 and spans[-1][1] == TOTAL and all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    out.put((rank, bool(ok), bool(cover)))
    dist.destroy_process_group()


def test_two_rank_sharding_and_stats_allreduce():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results == [(0, True, True), (1, True, True)]
