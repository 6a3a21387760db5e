"""Per-call cost of the gradient step's three all-reduce variants on N ranks (torchrun): f16_ppo_grad16 alone (no collective),
f16_ppo_grad16 + NCCL all_reduce + divide, f16_ppo_grad16_peer (exchange fused into the reduction kernel over peer memory).
CUDA graph of 32 calls each, CUDA events, max over ranks."""
import ctypes as C
import json
import math
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import f16_model_b200 as f16
from f16_model_b200 import _capi
from f16_model_b200.policy import pack_train_blob16
from f16_model_b200.sharding import PeerExchange

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
L = _capi.load_library()
n_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
N = 1 << 23
g = torch.Generator(device=dev).manual_seed(5 + rank)
pol = f16.TensorCorePolicy(device=dev, seed=3, logstd_init=-0.5)
rows = torch.rand((N, 8), device=dev, generator=g) * 2 - 1
rv = torch.rand((N, 2), device=dev, generator=g)
mb = torch.randperm(N, device=dev, generator=g)[:n_rows].contiguous()
blob = pack_train_blob16(pol.actor_mean, pol.critic)
scratch = torch.empty(L.f16_ppo_grad16_scratch_floats(local), device=dev)
stats = torch.tensor([0.0, 1.0], device=dev)
grad = torch.empty(L.f16_ppo_grad_floats(), device=dev)
px = PeerExchange.create(rank, world, dev)
args = (n_rows, blob.data_ptr(), rows.data_ptr(), None, None, None, rv.data_ptr(), None, mb.data_ptr(), pol.actor_logstd.data_ptr(),
        stats.data_ptr(), grad.data_ptr(), 0.2, 0.5, 0.01, 1, 1)
st = lambda: _capi.stream_ptr(dev)


def plain():
    _capi.check(L.f16_ppo_grad16(*args, scratch.data_ptr(), st()))


def nccl():
    plain()
    dist.all_reduce(grad)
    grad.div_(world)


def peer():
    _capi.check(L.f16_ppo_grad16_peer(*args, scratch.data_ptr(), C.byref(px.group), st()))


def timed(fn, reps=32, replays=8):
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(reps):
            fn()
    gr.replay()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e3 / (reps * replays)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    del gr
    return float(t.item())


out = {"ranks": world, "rows_per_rank": n_rows, "grad16_us": timed(plain), "grad16_nccl_div_us": timed(nccl)}
if px is not None:
    out["grad16_peer_us"] = timed(peer)
if rank == 0:
    print(json.dumps(out), flush=True)
torch.cuda.synchronize()
if px is not None:
    px.close()
dist.destroy_process_group()
