"""World-size-2 tests on two GPUs (NCCL over NVLink): the PPO gradient all-reduce of SURVEY 8e row 2.
Skipped on a box with fewer than two GPUs (run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py -m gpu`)."""
import math
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

CFG = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": False, "scenario": "combo"}


def _synthetic_rollout(f16, N, dev):
    """The same synthetic rollout rows on every process (seeded on the CPU generator, then moved)."""
    g = torch.Generator().manual_seed(11)
    pol = f16.TensorCorePolicy(device=dev, seed=3, logstd_init=-0.5)
    obs = (torch.rand((N, 5), generator=g) * 2 - 1).to(dev)
    with torch.no_grad():
        mean0, val0 = pol.actor_mean(obs).squeeze(-1), pol.critic(obs).squeeze(-1)
    std = math.exp(-0.5)
    actions = mean0 + std * torch.randn(N, generator=g).to(dev)
    logprobs = -0.5 * ((actions - mean0) / std) ** 2 + 0.5 - 0.5 * math.log(2 * math.pi) + 0.3 * torch.randn(N, generator=g).to(dev)
    adv = torch.randn(N, generator=g).to(dev) * 2 + 0.3
    ret = val0 + torch.randn(N, generator=g).to(dev)
    val = val0 + 0.4 * torch.randn(N, generator=g).to(dev)
    mb = torch.randperm(N, generator=g).to(dev)
    return pol, obs, actions, logprobs, adv, ret, val, mb


def _worker(rank, world, port, out):
    try:
        import torch.distributed as dist

        import f16_model_b200 as f16
        from f16_model_b200 import _capi
        from f16_model_b200.policy import pack_train_blob16
        from f16_model_b200.sharding import make_sharded_env, shard_range

        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        L = _capi.load_library()
        res = {}

        # (1) the gradient of a minibatch split over two ranks and averaged == the gradient of the whole minibatch on one rank
        N, n = 100000, 32768
        pol, obs, actions, logprobs, adv, ret, val, mb = _synthetic_rollout(f16, N, dev)
        blob = pack_train_blob16(pol.actor_mean, pol.critic)
        scratch = torch.empty(L.f16_ppo_grad16_scratch_floats(rank), device=dev)
        stats = torch.stack([adv[mb[:n]].mean(), adv[mb[:n]].std()])          # global minibatch statistics on both ranks

        def grad(idx):
            g = torch.empty(L.f16_ppo_grad_floats(), device=dev)
            _capi.check(L.f16_ppo_grad16(idx.numel(), blob.data_ptr(), obs.data_ptr(), actions.data_ptr(), logprobs.data_ptr(), adv.data_ptr(),
                                         ret.data_ptr(), val.data_ptr(), idx.data_ptr(), pol.actor_logstd.data_ptr(), stats.data_ptr(),
                                         g.data_ptr(), 0.2, 0.5, 0.01, 1, 1, scratch.data_ptr(), _capi.stream_ptr(dev)))
            return g

        whole = grad(mb[:n].contiguous())
        half = n // world
        mine = grad(mb[rank * half:(rank + 1) * half].contiguous())
        dist.all_reduce(mine)
        n_par = L.f16_ppo_grad_floats() - 5
        avg = mine[:n_par] / world                     # parameter gradients: each rank's kernel divides by its own row count
        rel = ((avg - whole[:n_par]).norm() / whole[:n_par].norm()).item()
        res["grad_rel_err"] = rel
        res["stats_rel_err"] = ((mine[n_par:] - whole[n_par:]).abs() / (whole[n_par:].abs() + 1.0)).max().item()   # the five sums add up

        # (1b) the same all-reduce fused into the gradient-reduction kernel over peer memory (f16_ppo_grad16_peer): the MEAN over the
        # ranks, bit-identical on both ranks and equal to the NCCL result; eager calls and CUDA-graph replays (slot reuse, counters)
        import ctypes as C

        from f16_model_b200.sharding import PeerExchange

        px = PeerExchange.create(rank, world, dev)
        res["peer_mapped"] = px is not None
        if px is not None:
            idx = mb[rank * half:(rank + 1) * half].contiguous()
            gp = torch.empty(L.f16_ppo_grad_floats(), device=dev)

            def grad_peer():
                _capi.check(L.f16_ppo_grad16_peer(idx.numel(), blob.data_ptr(), obs.data_ptr(), actions.data_ptr(), logprobs.data_ptr(),
                                                  adv.data_ptr(), ret.data_ptr(), val.data_ptr(), idx.data_ptr(), pol.actor_logstd.data_ptr(),
                                                  stats.data_ptr(), gp.data_ptr(), 0.2, 0.5, 0.01, 1, 1, scratch.data_ptr(), C.byref(px.group),
                                                  _capi.stream_ptr(dev)))

            errs = []
            for _ in range(5):
                gp.fill_(7.0)
                grad_peer()
                errs.append(((gp - mine / world).abs() / (mine.abs() / world + 1e-6)).max().item())
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                grad_peer()
            torch.cuda.current_stream(dev).wait_stream(side)
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for _ in range(3):
                    grad_peer()
            for _ in range(4):
                gp.fill_(7.0)
                gr.replay()
                errs.append(((gp - mine / world).abs() / (mine.abs() / world + 1e-6)).max().item())
            res["peer_rel_err"] = max(errs)
            other = gp.clone()
            dist.broadcast(other, 0)
            res["peer_ranks_equal"] = bool(torch.equal(other, gp))
            res["peer_calls"] = int(px.seq.min().item())
            torch.cuda.synchronize(dev)
            px.close()

        # (2) the trainer: two updates with the all-reduce inside the update graph and a KL early stop; parameters stay identical
        envs = make_sharded_env(CFG, 8192 * world, rank, world, dev, seed=1, fast_math=True)
        pol2 = f16.TensorCorePolicy(device=dev, seed=1, logstd_init=-1.0, env_id_base=shard_range(8192 * world, rank, world)[0])
        tr = f16.PPOTrainer(envs, pol2, f16.PPOConfig(num_steps=32, num_minibatches=8, update_epochs=4, learning_rate=1e-3, target_kl=0.02))
        hist = tr.train(num_updates=3)
        res["graph"] = bool(tr._update_graph)
        res["trainer_peer"] = tr._peer is not None
        flat = torch.cat([p.detach().reshape(-1) for p in pol2.parameters()])
        ref = flat.clone()
        dist.broadcast(ref, 0)
        res["params_equal"] = bool(torch.equal(ref, flat))
        res["finite"] = all(math.isfinite(h["value_loss"]) and math.isfinite(h["approx_kl"]) for h in hist)
        # ranks really hold different data: their rollouts differ
        r0 = tr.rewards.sum().reshape(1).clone()
        both = [torch.zeros_like(r0) for _ in range(world)]
        dist.all_gather(both, r0)
        res["different_shards"] = bool(both[0].item() != both[1].item())
        tr.close()
        # the same training with the NCCL all-reduce: same parameters up to the order of the fp32 atomics inside each rank's kernel
        envs_n = make_sharded_env(CFG, 8192 * world, rank, world, dev, seed=1, fast_math=True)
        pol_n = f16.TensorCorePolicy(device=dev, seed=1, logstd_init=-1.0, env_id_base=shard_range(8192 * world, rank, world)[0])
        tr_n = f16.PPOTrainer(envs_n, pol_n, f16.PPOConfig(num_steps=32, num_minibatches=8, update_epochs=4, learning_rate=1e-3, target_kl=0.02,
                                                           peer_allreduce=False))
        tr_n.train(num_updates=3)
        flat_n = torch.cat([p.detach().reshape(-1) for p in pol_n.parameters()])
        res["nccl_trainer_peer"] = tr_n._peer is not None
        res["peer_vs_nccl_params"] = ((flat_n - flat).norm() / flat.norm()).item()
        tr_n.close()
        # ... and with two launches per minibatch (f16_ppo_step16 with the peer exchange) instead of the persistent epoch kernel, whose reduction
        # phase makes the same exchange: same parameters up to the summation order of the gradient norm
        envs_s = make_sharded_env(CFG, 8192 * world, rank, world, dev, seed=1, fast_math=True)
        pol_s = f16.TensorCorePolicy(device=dev, seed=1, logstd_init=-1.0, env_id_base=shard_range(8192 * world, rank, world)[0])
        tr_s = f16.PPOTrainer(envs_s, pol_s, f16.PPOConfig(num_steps=32, num_minibatches=8, update_epochs=4, learning_rate=1e-3, target_kl=0.02,
                                                           fused_epoch=False))
        tr_s.train(num_updates=3)
        flat_s = torch.cat([p.detach().reshape(-1) for p in pol_s.parameters()])
        res["epoch_vs_step16_params"] = ((flat_s - flat).norm() / flat.norm()).item()
        res["step16_trainer_peer"] = tr_s._peer is not None
        tr_s.close()
        dist.destroy_process_group()
        out.put((rank, res))
    except Exception as exc:   # surface the failure instead of a silent timeout
        import traceback
        out.put((rank, {"error": f"{exc}\n{traceback.format_exc()}"}))


def test_two_rank_ppo_gradient_allreduce():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(out.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for r in (0, 1):
        res = results[r]
        assert "error" not in res, res.get("error")
        assert res["grad_rel_err"] < 1e-4, res           # fp32 sums in a different order
        assert res["stats_rel_err"] < 1e-4, res
        assert res["graph"] and res["params_equal"] and res["finite"] and res["different_shards"], res
        assert res["peer_mapped"] and res["trainer_peer"] and not res["nccl_trainer_peer"], res    # NVLink peer mapping works on a B200 box
        assert res["peer_rel_err"] < 1e-6 and res["peer_ranks_equal"] and res["peer_calls"] == 5 + 1 + 4 * 3, res
        assert res["peer_vs_nccl_params"] < 1e-3, res
        assert res["step16_trainer_peer"] and res["epoch_vs_step16_params"] < 1e-4, res   # epoch kernel (default) vs two launches per minibatch
