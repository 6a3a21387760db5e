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


def trim_grid_sharded(V, H, rank, world_size, device="cuda", group=None, trim_fn=None, **kw):
    """Batched trim search over a grid of flight conditions striped across ranks: every rank
    trims its contiguous block of (V, H) cells on its own GPU, then ONE ``all_gather`` of ``[cells, 5]``
    doubles (stab, throttle, alpha, cost, iterations) gives every rank the whole grid.  Returns the same dict
    as ``trim.trim_grid`` for all cells.  ``trim_fn`` (default ``trim.trim_grid``) does the per-rank work."""
    import numpy as np
    import torch
    import torch.distributed as dist

    if trim_fn is None:
        from .trim import trim_grid as trim_fn
    V = np.atleast_1d(np.asarray(V, dtype=np.float64)).ravel()
    H = np.atleast_1d(np.asarray(H, dtype=np.float64)).ravel()
    if V.shape != H.shape:
        raise ValueError("V and H must have the same number of cells")
    begin, end = shard_range(V.size, rank, world_size)
    n_max = (V.size + world_size - 1) // world_size          # blocks differ by at most one cell -> pad to the largest
    mine = np.zeros((n_max, 5))
    n_starts = 0
    if end > begin:
        r = trim_fn(V[begin:end], H[begin:end], device=device, **kw)
        mine[: end - begin] = np.stack([r["stab"], r["throttle"], r["alpha"], r["cost"], r["iterations"]], axis=1)
        n_starts = int(r.get("n_starts", 0))
    if world_size > 1:
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("trim_grid_sharded with world_size > 1 needs an initialised process group")
        on_gpu = dist.get_backend(group) == "nccl"
        t = torch.tensor(mine, device=device if on_gpu else "cpu")
        parts = [torch.empty_like(t) for _ in range(world_size)]
        dist.all_gather(parts, t, group=group)
        blocks = []
        for r_, part in enumerate(parts):
            b, e = shard_range(V.size, r_, world_size)
            blocks.append(part[: e - b].cpu().numpy())
        full = np.concatenate(blocks, axis=0)
    else:
        full = mine[: V.size]
    out = {"stab": full[:, 0], "throttle": full[:, 1], "alpha": full[:, 2], "cost": full[:, 3],
           "iterations": full[:, 4].astype(np.int32), "n_starts": n_starts}
    out["accepted"] = out["cost"] <= 10e-6
    return out


class PeerExchange:
    """The exchange buffers of ``f16_ppo_grad16_peer`` (the gradient all-reduce fused into the gradient-reduction kernel over
    NVLink / NVSwitch peer memory): every rank of the node allocates one buffer, the CUDA-IPC handles travel through ONE
    ``all_gather`` of 64 bytes per rank, every rank maps the others' buffers and keeps the pointer table on its device.
    ``group`` = the C struct to pass to the kernel launch.  Ranks must live on one node with peer access between their
    devices; ``PeerExchange.create`` returns None (and the trainer keeps the NCCL all-reduce) when the mapping fails on
    any rank.  Call ``close()`` before the process group is destroyed."""

    def __init__(self, rank, world, device, own, handles):
        import ctypes as C

        import torch

        from . import _capi

        L = _capi.load_library()
        self.rank, self.world, self.device = rank, world, device
        self._own, self._opened = own, []
        ptrs = []
        for r in range(world):
            if r == rank:
                ptrs.append(own)
                continue
            p = C.c_void_p()
            buf = (C.c_ubyte * 64).from_buffer_copy(bytes(handles[r]))
            _capi.check(L.f16_peer_open(buf, C.byref(p)))
            self._opened.append(p.value)
            ptrs.append(p.value)
        self.bufs = torch.tensor(ptrs, dtype=torch.int64, device=device)
        self.seq = torch.zeros(L.f16_peer_seq_words(), dtype=torch.int32, device=device)
        self.group = _capi.PeerGroup(rank, world, self.bufs.data_ptr(), self.seq.data_ptr())

    @classmethod
    def create(cls, rank, world, device, group=None):
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _capi

        L = _capi.load_library()
        ok, own, handle = 1, None, (C.c_ubyte * 64)()
        with torch.cuda.device(device):
            try:
                p = C.c_void_p()
                _capi.check(L.f16_peer_alloc(L.f16_peer_buffer_bytes(world), C.byref(p), handle))
                own = p.value
            except Exception:
                ok = 0
            on_gpu = dist.get_backend(group) == "nccl"
            dev = device if on_gpu else "cpu"
            mine = torch.tensor(list(bytes(handle)) + [ok], dtype=torch.uint8, device=dev)
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine, group=group)
            rows = [x.cpu().numpy() for x in parts]
            self = None
            if all(int(r[64]) == 1 for r in rows):
                try:
                    self = cls(rank, world, device, own, [r[:64].tobytes() for r in rows])
                except Exception:
                    self = None
            # every rank must agree: one failed mapping sends everybody back to NCCL
            flag = torch.tensor([1 if self is not None else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 0:
                if self is not None:
                    self.close()
                elif own is not None:
                    L.f16_peer_free(own)
                return None
            return self

    def close(self):
        from . import _capi

        L = _capi.load_library()
        import torch

        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for p in self._opened:
                L.f16_peer_close(p)
            self._opened = []
            if self._own is not None:
                L.f16_peer_free(self._own)
                self._own = None
