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

from f16_model_b200.sharding import global_episode_stats, shard_range, trim_grid_sharded

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
    # trim grid striped over ranks + one all_gather: every rank ends up with the whole grid, cell for cell
    Vg, Hg = np.meshgrid(np.linspace(100, 300, 7), np.linspace(1000, 10000, 5))     # 35 cells: ragged over 2 ranks
    seen = []

    def fake_trim(V, H, device=None, **kw):     # stands in for the CUDA search: a closed form of the cell
        seen.append(len(V))
        return {"stab": -V / 1000, "throttle": H / 20000, "alpha": V / H, "cost": 1e-7 * V, "iterations": (V + H).astype(np.int32),
                "n_starts": 11}

    g = trim_grid_sharded(Vg, Hg, rank, world, device="cpu", trim_fn=fake_trim)
    Vf, Hf = Vg.ravel(), Hg.ravel()
    b, e = shard_range(Vf.size, rank, world)
    trim_ok = (seen == [e - b] and np.array_equal(g["stab"], -Vf / 1000) and np.array_equal(g["throttle"], Hf / 20000)
               and np.array_equal(g["alpha"], Vf / Hf) and np.array_equal(g["iterations"], (Vf + Hf).astype(np.int32))
               and np.array_equal(g["accepted"], 1e-7 * Vf <= 1e-5) and g["n_starts"] == 11)
    out.put((rank, bool(ok), bool(cover and trim_ok)))
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
