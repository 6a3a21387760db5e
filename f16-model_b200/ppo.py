"""On-device PPO -- mirror of ``Agent.train`` (control/ppo/ppo_model_gsde.py:181-398) with the
host round trips removed: the rollout loop is ``tensor-core policy -> fused env-step kernel``
on device tensors, GAE is one CUDA kernel, the clipped-PPO update uses torch autograd on the
same parameters and Adam(amsgrad), and with more than one rank the flat gradient (~9.3 k floats)
is all-reduced over NCCL once per minibatch.

Flags keep the reference's names and defaults (control/ppo/utils.py:11-142).  Deliberate
deviations: the reference's gSDE noise path is broken as committed
(``theta_gsde`` is never created in ``train``); here exploration is the standard Gaussian policy
``a = mean + exp(logstd) * eps`` with eps drawn by the counter RNG inside the policy kernel."""
import ctypes as C
import dataclasses
import time

from . import _capi


@dataclasses.dataclass
class PPOConfig:
    learning_rate: float = 3e-4
    seed: int = 1
    total_timesteps: int = 200000
    num_steps: int = 2048
    anneal_lr: bool = True
    gae: bool = True
    gamma: float = 0.99
    gae_lambda: float = 0.95
    num_minibatches: int = 64
    update_epochs: int = 10
    norm_adv: bool = True
    clip_coef: float = 0.5
    clip_vloss: bool = True
    ent_coef: float = 0.0
    vf_coef: float = 0.5
    max_grad_norm: float = 0.5
    target_kl: float = None
    reset_every_update: bool = True      # the reference re-resets the envs at every update (:210)
    use_cuda_graph: bool = True          # replay the whole rollout (2 kernels per step) from one CUDA graph
    tf32_update: bool = False            # run the update's matmuls on the TF32 tensor cores (default: torch's fp32)
    fused_update: bool = True            # minibatch forward + loss + backward in ONE tcgen05 kernel (f16_ppo_grad16); False: torch autograd
    fused_precision: str = "fp16"        # tensor-core operands of the fused kernel: "fp16" (weight gradients on the tensor cores too) or "tf32"
    fused_optimizer: bool = True         # fused fp16 update: reduction (+ all-reduce) + clip + Adam + blob re-packing in ONE multi-CTA kernel (f16_ppo_step16)
                                         # instead of the reduction kernel followed by the single-CTA optimiser kernel
    fused_epoch: bool = True             # fused fp16 update (one rank, or ranks with the peer-memory exchange): ALL minibatch steps of an epoch in ONE persistent cooperative kernel (f16_ppo_epoch16:
                                         # the optimiser step is taken inside the gradient kernel, between grid-wide meetings) instead of two launches per minibatch
    peer_allreduce: bool = True          # world > 1, fused fp16 update: the gradient all-reduce rides in the gradient-reduction kernel over NVLink peer
                                         # memory (f16_ppo_grad16_peer) instead of an NCCL call per minibatch; falls back to NCCL when the ranks cannot map each other
    graph_allreduce: bool = True         # world > 1: capture the NCCL gradient all-reduce inside the update's CUDA graph (call close() before destroy_process_group)
    fused_rollout: bool = True           # the whole closed-loop rollout in ONE persistent kernel (f16_rollout): policy -> head -> env step, K times


def gae(rewards, values, done_after, next_value, gamma, gae_lambda):
    """Advantages and returns [T, N] by the CUDA reverse scan (f16_gae).  ``done_after`` is the uint8
    [T, N] done flag returned by each step (``dones[t+1]`` / ``next_done`` in the reference's loop)."""
    import torch

    T, N = rewards.shape
    adv = torch.empty_like(rewards)
    ret = torch.empty_like(rewards)
    L = _capi.load_library()
    for t in (rewards, values, next_value):
        _capi.require_cuda(t, "gae input")
        if t.dtype != torch.float32:
            raise ValueError("gae inputs must be float32")
    _capi.require_cuda(done_after, "done_after")
    if done_after.dtype not in (torch.uint8, torch.bool):
        raise ValueError("done_after must be uint8 / bool")
    with torch.cuda.device(rewards.device):
        _capi.check(L.f16_gae(T, N, rewards.data_ptr(), values.data_ptr(), done_after.data_ptr(), next_value.data_ptr(),
                              float(gamma), float(gae_lambda), adv.data_ptr(), ret.data_ptr(), _capi.stream_ptr(rewards.device)))
    return adv, ret


def fused_rollout(envs, blob16, logstd, obs0, obs_next, actions, logprobs, values, rewards, dones, next_value=None,
                  sample=True, policy_seed=0):
    """K closed-loop steps of ``policy -> Gaussian head -> F16.step`` for every env of ``envs`` in ONE kernel launch
    (csrc/f16_rollout.cu; the loop of ppo_model_gsde.py:216-230).  ``blob16`` = ``policy.pack_train_blob16(actor, critic)``
    (the PPO trainer's fp16 weight blob); ``obs0`` [N,5] is the observation the first step acts on; the [K,N(,5)] rollout
    rows and the bootstrap value of the last observation are written into the caller's tensors.  The env's episode
    statistics and ``final_*`` buffers are updated as by ``step``."""
    import ctypes as C
    import torch

    K, N = actions.shape
    if N != envs.num_envs or obs_next.shape != (K, N, 5) or obs0.shape != (N, 5):
        raise ValueError("rollout buffers must be [K, N] / [K, N, 5] with N = envs.num_envs, obs0 [N, 5]")
    for name, t in (("obs0", obs0), ("obs_next", obs_next), ("actions", actions), ("logprobs", logprobs), ("values", values),
                    ("rewards", rewards), ("blob16", blob16), ("logstd", logstd)):
        _capi.require_cuda(t, name)
        if t.dtype != torch.float32:
            raise ValueError(f"{name} must be float32")
    _capi.require_cuda(dones, "dones")
    if dones.dtype not in (torch.uint8, torch.bool) or dones.shape != (K, N):
        raise ValueError("dones must be uint8 [K, N]")
    io = _capi.RolloutIO()
    io.train_blob16, io.logstd, io.obs0, io.obs_next = blob16.data_ptr(), logstd.data_ptr(), obs0.data_ptr(), obs_next.data_ptr()
    io.actions, io.logprobs, io.values, io.rewards = actions.data_ptr(), logprobs.data_ptr(), values.data_ptr(), rewards.data_ptr()
    io.dones = dones.data_ptr()
    io.next_value = _capi.require_cuda(next_value, "next_value").data_ptr() if next_value is not None else None
    io.final_obs, io.final_ep_len, io.final_return = envs._final_obs.data_ptr(), envs._final_len.data_ptr(), envs._final_ret.data_ptr()
    io.stats = envs.episode_stats.data_ptr()
    io.sample = int(bool(sample))
    io.policy_seed = int(policy_seed) & 0xFFFFFFFFFFFFFFFF
    with torch.cuda.device(envs.device):
        _capi.check(_capi.load_library().f16_rollout(C.byref(envs._cfg), C.byref(envs._buf), C.byref(io), K, _capi.stream_ptr(envs.device)))
    envs._obs.copy_(obs_next[K - 1])          # the env's own "last observation" buffer stays current for step() users


def kl_early_stop(approx_kl, target_kl, world=1, group=None):
    """The early-stop test of the update loop (ppo_model_gsde.py:367-369), made rank-consistent: with more than one
    rank the approximate KL of the last minibatch is averaged over ranks before it is compared with ``target_kl``, so
    every rank leaves the epoch loop at the same epoch (a rank-local decision would leave the gradient all-reduce of
    the next epoch with mismatched participants -- NCCL would hang)."""
    if target_kl is None:
        return False
    import torch

    kl = approx_kl.detach().reshape(1).clone()
    if world > 1:
        torch.distributed.all_reduce(kl, group=group)
        kl /= world
    return bool(kl.item() > target_kl)


class PPOTrainer:
    def __init__(self, envs, policy, config=None):
        import torch

        self._torch = torch
        self.envs, self.policy, self.cfg = envs, policy, config or PPOConfig()
        self.N = envs.num_envs
        self.device = envs.device
        self.batch_size = self.N * self.cfg.num_steps
        self.minibatch_size = self.batch_size // self.cfg.num_minibatches
        # lr as a device tensor + capturable=True: the optimiser step can live inside a CUDA graph and still be annealed
        self._lr = torch.tensor(float(self.cfg.learning_rate), device=envs.device)
        self.optimizer = torch.optim.Adam(policy.parameters(), lr=self._lr, amsgrad=True, capturable=True)
        T, N, dev = self.cfg.num_steps, self.N, self.device
        aos = envs.obs_layout == "aos"
        self.obs = torch.zeros((T + 1, N, 5) if aos else (T + 1, 5, N), device=dev)   # obs[t] feeds step t; obs[T] bootstraps
        self.actions = torch.zeros((T, N), device=dev)
        self.logprobs = torch.zeros((T, N), device=dev)
        self.rewards = torch.zeros((T, N), device=dev)
        self.done_after = torch.zeros((T, N), dtype=torch.uint8, device=dev)            # done flag returned by step t
        self.values = torch.zeros((T, N), device=dev)
        self.global_step = 0
        self.update_idx = 0
        self.world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size()
        self._gen = torch.Generator(device=dev).manual_seed(self.cfg.seed)
        self._perm_calls = 0
        self._graph = None
        self._update_graph = None
        self._flat = None
        self._fused = bool(self.cfg.fused_update) and envs.obs_layout == "aos"
        self._pflat_seen = None
        self._peer = None
        if self._fused:
            self._setup_fused()
            if self.world > 1 and self._fp16 and self.cfg.peer_allreduce:
                from .sharding import PeerExchange

                self._peer = PeerExchange.create(torch.distributed.get_rank(), self.world, dev)   # None: the ranks cannot map each other -> NCCL
        self._next_value = torch.zeros(N, device=dev)
        c = envs._cfg
        self._fused_rollout = (bool(self.cfg.fused_rollout) and aos and envs.dtype == torch.float32 and envs.autoreset
                               and not c.noise and envs.pos_comp is None)
        self._rollout_blob = None

    # ------------------------------------------------------------------ rollout
    def _obs_at(self, t):
        return self.obs[t] if self.envs.obs_layout == "aos" else self.obs[t].t()

    def _enqueue_rollout(self):
        """2 kernels per step, no host work in between: the policy kernel writes action / log-prob / value
        straight into row t of the rollout buffers, the env kernel reads the action row and writes
        reward / done rows and the NEXT observation row."""
        for t in range(self.cfg.num_steps):
            self.policy.act(self._obs_at(t), sample=True, env=self.envs,
                            out=(self.actions[t], self.logprobs[t], self.values[t]))
            self.envs.step_k(self.actions[t:t + 1], out=(self.obs[t + 1], self.rewards[t:t + 1], self.done_after[t:t + 1]))

    def _blob16(self):
        """The fp16 weight blob the rollout kernel reads: the fused update's own training blob (kept current by the
        optimiser kernel), otherwise packed from the torch parameters."""
        if self._fused and getattr(self, "_fp16", False):
            return self._train_blob
        from .policy import pack_train_blob16

        blob = pack_train_blob16(self.policy.actor_mean, self.policy.critic)
        if self._rollout_blob is None:
            self._rollout_blob = blob
        else:
            self._rollout_blob.copy_(blob)
        return self._rollout_blob

    def rollout(self):
        """num_steps closed-loop steps for all envs, everything device-resident: ONE persistent kernel
        (``fused_rollout``, the default), or 2 kernels per step replayed from a CUDA graph."""
        torch, cfg = self._torch, self.cfg
        if cfg.reset_every_update or self.update_idx == 0:
            obs0, _ = self.envs.reset(seed=cfg.seed + self.update_idx + 1)
            self._obs_at(0).copy_(obs0)
        else:
            self.obs[0].copy_(self.obs[cfg.num_steps])
        if self._fused_rollout:
            T = cfg.num_steps
            fused_rollout(self.envs, self._blob16(), self.policy.actor_logstd.detach(), self.obs[0], self.obs[1:], self.actions,
                          self.logprobs, self.values, self.rewards, self.done_after, next_value=self._next_value,
                          sample=True, policy_seed=self.policy.seed)
            self.global_step += self.N * T
            adv, ret = gae(self.rewards, self.values, self.done_after, self._next_value, cfg.gamma, cfg.gae_lambda)
            s = self.envs.episode_stats.clone()
            self.envs.episode_stats.zero_()
            return adv, ret, s[1], s[0]
        if not cfg.use_cuda_graph:
            self._enqueue_rollout()
        else:
            if self._graph is None:
                side = torch.cuda.Stream(self.device)
                side.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(side):                      # warm-up outside capture
                    self.policy.act(self._obs_at(0), sample=True, env=self.envs,
                                    out=(self.actions[0], self.logprobs[0], self.values[0]))
                torch.cuda.current_stream(self.device).wait_stream(side)
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph):
                    self._enqueue_rollout()
            self._graph.replay()
        self.global_step += self.N * cfg.num_steps
        _, next_value = self.policy.mean_value(self._obs_at(cfg.num_steps))
        adv, ret = gae(self.rewards, self.values, self.done_after, next_value, cfg.gamma, cfg.gae_lambda)
        # episode statistics come from the env kernel's own accumulators (warp-level reductions)
        s = self.envs.episode_stats.clone()
        self.envs.episode_stats.zero_()
        ep_count, ep_returns = s[0], s[1]
        return adv, ret, ep_returns, ep_count

    # ------------------------------------------------------------------ update
    def _allreduce_grads(self):
        if self.world == 1:
            return
        torch = self._torch
        params = [p for p in self.policy.parameters() if p.grad is not None]
        flat = torch.cat([p.grad.reshape(-1) for p in params])
        torch.distributed.all_reduce(flat)                      # one NCCL call, ~37 KB: latency-bound
        flat /= self.world
        off = 0
        for p in params:
            p.grad.copy_(flat[off:off + p.numel()].view_as(p.grad))
            off += p.numel()

    # ---- fused gradient path -------------------------------------------------------------------
    def _setup_fused(self):
        """Persistent pieces of the fused path.  Parameters and gradients become views of two flat vectors in the
        kernel's field order (``PolicyGrad``): ``f16_ppo_grad`` accumulates into the gradient vector, ``f16_ppo_adam``
        clips it, applies Adam(amsgrad) to the parameter vector and re-packs both weight blobs (blob = flat[perm])."""
        torch = self._torch
        from .policy import grad_views, pack_policy_blob, pack_train_blob, pack_train_blob16
        from torch import nn

        if self.cfg.fused_precision not in ("fp16", "tf32"):
            raise ValueError("fused_precision must be 'fp16' or 'tf32'")
        self._fp16 = self.cfg.fused_precision == "fp16"

        L = _capi.load_library()
        pol, dev = self.policy, self.device
        n_all = L.f16_ppo_grad_floats()
        self._n_params = n_all - 5                                 # the last five floats are loss statistics
        self._g = torch.zeros(n_all, device=dev)
        self._pflat = torch.zeros(n_all, device=dev)
        G, W = grad_views(self._g), grad_views(self._pflat)
        la = [m for m in pol.actor_mean if isinstance(m, nn.Linear)]
        lc = [m for m in pol.critic if isinstance(m, nn.Linear)]
        with torch.no_grad():
            for net, ls in enumerate((la, lc)):
                for layer, (wk, bk) in enumerate((("w1", "b1"), ("w2", "b2"), ("w3", "b3"))):
                    wv = W[wk][net].view_as(ls[layer].weight)
                    bv = W[bk][net:net + 1].view_as(ls[layer].bias) if bk == "b3" else W[bk][net]
                    wv.copy_(ls[layer].weight)
                    bv.copy_(ls[layer].bias)
                    ls[layer].weight.data, ls[layer].bias.data = wv, bv
                    ls[layer].weight.grad = G[wk][net].view_as(wv)
                    ls[layer].bias.grad = G[bk][net:net + 1].view_as(bv) if bk == "b3" else G[bk][net]
            W["logstd"].copy_(pol.actor_logstd)
            pol.actor_logstd.data = W["logstd"]
            pol.actor_logstd.grad = G["logstd"]
            # blob permutations: pack both blobs once from parameters holding (their flat index + 1); 0 marks padding
            keep = self._pflat.clone()
            self._pflat[: self._n_params] = torch.arange(1, self._n_params + 1, device=dev, dtype=torch.float32)
            self._fwd_perm = pack_policy_blob(pol.actor_mean, pol.critic).round().long()
            if self._fp16:
                halves, tail = pack_train_blob16(pol.actor_mean, pol.critic, logical=True)
                self._half_perm, self._train_perm = halves.round().long(), tail.round().long()
            else:
                self._half_perm = None
                self._train_perm = pack_train_blob(pol.actor_mean, pol.critic).round().long()
            self._pflat.copy_(keep)
        self._gstats = G["stats"]
        self._train_blob = (pack_train_blob16 if self._fp16 else pack_train_blob)(pol.actor_mean, pol.critic)
        # fp16 blob: [fp16 tensor-core operands | fp32 biases and heads]; the optimiser kernel writes the two regions separately
        self._blob_tail = self._train_blob[self._half_perm.numel() // 2:] if self._fp16 else self._train_blob
        if self._fp16:
            pol._blob16 = self._train_blob     # policy.act(precision="fp16") reads the blob the optimiser kernel keeps current
        self._step16 = None
        if self._fp16 and self.cfg.fused_optimizer:
            from .policy import scatter_table

            self._scatter = scatter_table(self._n_params, self._fwd_perm, self._train_perm, self._half_perm)
            self._sync16 = torch.zeros(L.f16_ppo_step16_sync_bytes(), dtype=torch.uint8, device=dev)
            self._sync_epoch = torch.zeros(L.f16_ppo_epoch16_sync_bytes(), dtype=torch.uint8, device=dev)
            self._step16 = _capi.PpoStep16()
        self._params = list(pol.parameters())
        self._adv_stats = torch.zeros(2, device=dev)
        self._adv_scratch = torch.zeros(4, device=dev, dtype=torch.float64)
        self._grad_scratch = torch.empty(max(1, L.f16_ppo_grad16_scratch_floats(dev.index)), device=dev) if self._fp16 else None
        self._adam = [torch.zeros(self._n_params, device=dev) for _ in range(3)] + [torch.zeros(1, device=dev)]   # m, v, vmax, step

    def _fused_minibatch_step(self, mb, adv_stats=None, epoch_minibatches=0):
        """The same optimiser step in two kernels: forward + loss + backward (ppo_grad_kernel), then clip + Adam + blob
        re-packing (ppo_adam_kernel); csrc/f16_policy.cu.  With ``epoch_minibatches`` = k > 0 (single rank, fused fp16 optimiser): the k
        consecutive minibatch steps whose indices start at ``mb`` and whose advantage statistics start at ``adv_stats``, in one
        persistent kernel (f16_ppo_epoch16)."""
        torch, cfg = self._torch, self.cfg
        b = self._flat
        L = _capi.load_library()
        m, v, vmax, step = self._adam                      # (the gradient buffer is cleared by the previous optimiser kernel)
        with torch.cuda.device(self.device):
            st = _capi.stream_ptr(self.device)
            if adv_stats is None:
                adv_stats = self._adv_stats
                if cfg.norm_adv:
                    _capi.check(L.f16_ppo_adv_stats(mb.numel(), 1, b["adv"].data_ptr(), mb.data_ptr(), adv_stats.data_ptr(),
                                                    self._adv_scratch.data_ptr(), st))
            if self._fp16:   # packed rows (f16_ppo_pack_rows, once per update): one DRAM sector per gathered row
                srcs = (self._rows.data_ptr(), None, None, None, self._rv.data_ptr(), None)
            else:
                srcs = (b["obs"].data_ptr(), b["actions"].data_ptr(), b["logprobs"].data_ptr(), b["adv"].data_ptr(), b["ret"].data_ptr(),
                        b["val"].data_ptr())
            args = (mb.numel(), self._train_blob.data_ptr(), *srcs,
                    mb.data_ptr(), self.policy.actor_logstd.data_ptr(), adv_stats.data_ptr(),
                    self._g.data_ptr(), float(cfg.clip_coef), float(cfg.vf_coef), float(cfg.ent_coef),
                    int(bool(cfg.norm_adv)), int(bool(cfg.clip_vloss)))
            peer = self._peer if self._fp16 else None
            if epoch_minibatches:   # every minibatch step of the epoch in ONE persistent cooperative kernel; `mb` is the first minibatch's slice of the permutation
                assert self._step16 is not None and (self.world == 1 or peer is not None)
            if self._step16 is not None and (self.world == 1 or peer is not None):   # gradient kernel + ONE kernel for reduction / all-reduce / clip / Adam / blob scatter
                a = self._step16
                (a.n, a.train_blob16, a.obs, a.actions, a.logprobs, a.adv, a.ret, a.val, a.mb_idx, a.logstd, a.adv_stats, a.grad, a.clip_coef,
                 a.vf_coef, a.ent_coef, a.norm_adv, a.clip_vloss) = args
                a.scratch = self._grad_scratch.data_ptr()
                a.peer = C.pointer(peer.group) if peer is not None else None
                a.n_params = self._n_params
                a.param, a.exp_avg, a.exp_avg_sq, a.max_exp_avg_sq, a.step = (self._pflat.data_ptr(), m.data_ptr(), v.data_ptr(),
                                                                                vmax.data_ptr(), step.data_ptr())
                a.lr, a.beta1, a.beta2, a.eps, a.max_grad_norm = self._lr.data_ptr(), 0.9, 0.999, 1e-8, float(cfg.max_grad_norm)
                a.scatter, a.fwd_blob, a.tail_blob, a.half_blob = (self._scatter.data_ptr(), self.policy._blob.data_ptr(),
                                                                    self._blob_tail.data_ptr(), self._train_blob.data_ptr())
                a.diag, a.inv_rows, a.sync = self._diag.data_ptr(), 1.0 / mb.numel(), self._sync16.data_ptr()
                if epoch_minibatches:
                    a.sync = self._sync_epoch.data_ptr()
                    _capi.check(L.f16_ppo_epoch16(C.byref(a), int(epoch_minibatches), st))
                else:
                    _capi.check(L.f16_ppo_step16(C.byref(a), st))
                return
            if peer is not None:   # ... and the cross-rank mean in the same reduction kernel (peer memory, no NCCL call)
                _capi.check(L.f16_ppo_grad16_peer(*args, self._grad_scratch.data_ptr(), C.byref(peer.group), st))
            elif self._fp16:   # per-CTA partial gradients + a reduction kernel instead of 1.4 M atomics
                _capi.check(L.f16_ppo_grad16(*args, self._grad_scratch.data_ptr(), st))
            else:
                _capi.check(L.f16_ppo_grad(*args, st))
            if self.world > 1 and peer is None:
                torch.distributed.all_reduce(self._g)              # one NCCL call on the flat buffer (gradients and statistics), captured
                self._g /= self.world                              # in the update graph like the kernels around it
            _capi.check(L.f16_ppo_adam(self._n_params, self._pflat.data_ptr(), self._g.data_ptr(), m.data_ptr(), v.data_ptr(),
                                       vmax.data_ptr(), step.data_ptr(), self._lr.data_ptr(), 0.9, 0.999, 1e-8,
                                       float(cfg.max_grad_norm), self._fwd_perm.data_ptr(), self.policy._blob.data_ptr(),
                                       self._fwd_perm.numel(), self._train_perm.data_ptr(), self._blob_tail.data_ptr(),
                                       self._train_perm.numel(), self._half_perm.data_ptr() if self._fp16 else None,
                                       self._train_blob.data_ptr() if self._fp16 else None,
                                       self._half_perm.numel() if self._fp16 else 0, self._diag.data_ptr(),
                                       self.policy.actor_logstd.data_ptr(), 1.0 / mb.numel(), 1, st))

    def _state_tensors(self):
        """Everything a fused optimiser step mutates (snapshotted around the CUDA-graph warm-up)."""
        return [self._pflat, *self._adam, self._train_blob, self.policy._blob]

    def _minibatch_step(self, mb):
        """One clipped-PPO optimiser step on the minibatch indices ``mb`` (ppo_model_gsde.py:303-364).
        Writes its diagnostics into persistent scalars so that it can be replayed from a CUDA graph."""
        torch, cfg = self._torch, self.cfg
        if self._fused:
            return self._fused_minibatch_step(mb)
        b = self._flat
        newlogprob, entropy, newvalue = self.policy.evaluate(b["obs"][mb], b["actions"][mb])
        logratio = newlogprob - b["logprobs"][mb]
        ratio = logratio.exp()
        with torch.no_grad():
            self._diag[0].copy_((-logratio).mean())                                   # old_approx_kl
            self._diag[1].copy_(((ratio - 1) - logratio).mean())                      # approx_kl
            self._diag[2].add_(((ratio - 1.0).abs() > cfg.clip_coef).float().mean())  # clip fraction, summed
        mb_adv = b["adv"][mb]
        if cfg.norm_adv:
            mb_adv = (mb_adv - mb_adv.mean()) / (mb_adv.std() + 1e-8)
        pg_loss = torch.max(-mb_adv * ratio, -mb_adv * torch.clamp(ratio, 1 - cfg.clip_coef, 1 + cfg.clip_coef)).mean()
        if cfg.clip_vloss:
            v_unclipped = (newvalue - b["ret"][mb]) ** 2
            v_clipped = b["val"][mb] + torch.clamp(newvalue - b["val"][mb], -cfg.clip_coef, cfg.clip_coef)
            v_loss = 0.5 * torch.max(v_unclipped, (v_clipped - b["ret"][mb]) ** 2).mean()
        else:
            v_loss = 0.5 * ((newvalue - b["ret"][mb]) ** 2).mean()
        entropy_loss = entropy.mean()
        loss = pg_loss - cfg.ent_coef * entropy_loss + v_loss * cfg.vf_coef
        self.optimizer.zero_grad(set_to_none=False)
        loss.backward()
        self._allreduce_grads()
        torch.nn.utils.clip_grad_norm_(self.policy.parameters(), cfg.max_grad_norm)
        self.optimizer.step()
        with torch.no_grad():
            self._diag[3].copy_(v_loss)
            self._diag[4].copy_(pg_loss)
            self._diag[5].copy_(entropy_loss)

    def _prepare_update(self, advantages, returns):
        torch, cfg = self._torch, self.cfg
        T = cfg.num_steps
        if self._flat is None:
            self._adv = torch.empty_like(advantages)
            self._ret = torch.empty_like(returns)
            self._flat = {
                "obs": (self.obs[:T] if self.envs.obs_layout == "aos" else self.obs[:T].transpose(1, 2)).reshape(-1, 5),
                "logprobs": self.logprobs.reshape(-1), "actions": self.actions.reshape(-1),
                "adv": self._adv.reshape(-1), "ret": self._ret.reshape(-1), "val": self.values.reshape(-1)}
            self._diag = torch.zeros(6, device=self.device)
            self._mb_idx = torch.zeros(self.minibatch_size, dtype=torch.long, device=self.device)
            self._perm_buf = torch.arange(self.batch_size, dtype=torch.long, device=self.device)
            n_mb = max(1, self.batch_size // self.minibatch_size)
            self._adv_stats_epoch = torch.zeros((n_mb, 2), device=self.device)
            self._adv_scratch_epoch = torch.zeros((n_mb, 4), device=self.device, dtype=torch.float64)
        self._adv.copy_(advantages)
        self._ret.copy_(returns)
        self._diag.zero_()
        if self._fused and self._fp16:
            b = self._flat
            if getattr(self, "_rows", None) is None:
                self._rows = torch.empty((self.batch_size, 8), device=self.device)
                self._rv = torch.empty((self.batch_size, 2), device=self.device)
            with torch.cuda.device(self.device):
                _capi.check(_capi.load_library().f16_ppo_pack_rows(
                    self.batch_size, b["obs"].data_ptr(), b["actions"].data_ptr(), b["logprobs"].data_ptr(), b["adv"].data_ptr(),
                    b["ret"].data_ptr(), b["val"].data_ptr(), self._rows.data_ptr(), self._rv.data_ptr(), _capi.stream_ptr(self.device)))

    def _capture_update(self):
        """Capture one optimiser step (forward, backward, clip, Adam) in a CUDA graph: ~150 small launches
        per minibatch become one."""
        torch = self._torch
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        params = list(self.policy.parameters())
        if self._fused:
            state = self._state_tensors()
            saved = [t.detach().clone() for t in state]
        else:                                              # torch creates the Adam state lazily: absent before the first step
            saved = [p.detach().clone() for p in params]
            saved_opt = [{k: v.clone() for k, v in self.optimizer.state.get(p, {}).items() if torch.is_tensor(v)} for p in params]
        with torch.cuda.stream(side):                      # warm-up iterations (allocator, autograd, optimiser state)
            for _ in range(3):
                self._minibatch_step(self._mb_idx)
        torch.cuda.current_stream(self.device).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            if self._fused:
                # a whole epoch in one graph: every minibatch reads its slice of the persistent permutation buffer, so an epoch
                # costs one randperm and one graph launch (4 kernels per minibatch)
                starts = range(0, self.batch_size - self.minibatch_size + 1, self.minibatch_size)
                if self.cfg.norm_adv:                      # the advantage statistics of all the epoch's minibatches in one launch
                    with torch.cuda.device(self.device):
                        _capi.check(_capi.load_library().f16_ppo_adv_stats(
                            self.minibatch_size, len(starts), self._flat["adv"].data_ptr(), self._perm_buf.data_ptr(),
                            self._adv_stats_epoch.data_ptr(), self._adv_scratch_epoch.data_ptr(), _capi.stream_ptr(self.device)))
                if self._step16 is not None and (self.world == 1 or self._peer is not None) and self.cfg.fused_epoch:
                    self._fused_minibatch_step(self._perm_buf[: self.minibatch_size], adv_stats=self._adv_stats_epoch[0], epoch_minibatches=len(starts))
                else:
                    for k, start in enumerate(starts):
                        self._fused_minibatch_step(self._perm_buf[start:start + self.minibatch_size], adv_stats=self._adv_stats_epoch[k])
            else:
                self._minibatch_step(self._mb_idx)
        with torch.no_grad():                              # undo the warm-up steps IN PLACE (the graph holds these tensors)
            if self._fused:
                for t, q in zip(state, saved):
                    t.copy_(q)
            else:
                for p, q, so in zip(params, saved, saved_opt):
                    p.copy_(q)
                    for k, v in self.optimizer.state[p].items():
                        if torch.is_tensor(v):
                            v.copy_(so[k]) if k in so else v.zero_()
        self._diag.zero_()
        return graph

    def update(self, advantages, returns):
        torch, cfg = self._torch, self.cfg
        prev_tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = bool(cfg.tf32_update)
        try:
            return self._update(advantages, returns)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev_tf32

    def _update(self, advantages, returns):
        torch, cfg = self._torch, self.cfg
        self._prepare_update(advantages, returns)
        if self._fused:
            self._g.zero_()
            # the optimiser kernels keep the weight blobs current; they are re-packed from the torch parameters only when those were changed
            # from outside since the last update (load_agent, a manual edit): one compare instead of ~40 small packing launches per update
            if self._pflat_seen is None or not torch.equal(self._pflat_seen, self._pflat):
                from .policy import pack_train_blob, pack_train_blob16
                self._train_blob.copy_((pack_train_blob16 if self._fp16 else pack_train_blob)(self.policy.actor_mean, self.policy.critic))
                self.policy.sync_weights()
        use_graph = cfg.use_cuda_graph and (self.world == 1 or cfg.graph_allreduce)
        if use_graph and self._update_graph is None:
            try:
                self._update_graph = self._capture_update()
            except Exception as exc:                       # capture is an optimisation; fall back to eager launches
                print(f"[ppo] update graph capture failed ({exc}); running eagerly")
                self._update_graph = False
        n_steps = 0
        per_epoch = len(range(0, self.batch_size - self.minibatch_size + 1, self.minibatch_size))
        for epoch in range(cfg.update_epochs):
            if self._fused:
                # the epoch's permutation from the keyed-bijection kernel (no sort: torch.randperm was a fifth of the fused update);
                # the same on every rank, as the reference's single process would draw it
                self._perm_calls += 1
                with torch.cuda.device(self.device):
                    _capi.check(_capi.load_library().f16_ppo_randperm(
                        self.batch_size, (int(cfg.seed) * 0x9E3779B97F4A7C15 + self._perm_calls) & 0xFFFFFFFFFFFFFFFF,
                        self._perm_buf.data_ptr(), _capi.stream_ptr(self.device)))
                perm = self._perm_buf
                if use_graph and self._update_graph:
                    self._update_graph.replay()
                    n_steps += per_epoch
                    if kl_early_stop(self._diag[1], cfg.target_kl, self.world):
                        break
                    continue
            else:
                perm = torch.randperm(self.batch_size, device=self.device, generator=self._gen)
            for start in range(0, self.batch_size - self.minibatch_size + 1, self.minibatch_size):
                mb = perm[start:start + self.minibatch_size]
                if use_graph and self._update_graph:
                    self._mb_idx.copy_(mb)
                    self._update_graph.replay()
                else:
                    self._minibatch_step(mb)
                n_steps += 1
            if kl_early_stop(self._diag[1], cfg.target_kl, self.world):
                break
        if self._fused:
            if self._pflat_seen is None:
                self._pflat_seen = self._pflat.clone()
            else:
                self._pflat_seen.copy_(self._pflat)                 # (the optimiser kernels re-packed the blobs with every step)
        else:
            self.policy.sync_weights()                              # torch optimiser: refresh the tensor-core weight blobs
        d = self._diag.tolist()
        return dict(old_approx_kl=d[0], approx_kl=d[1], clipfrac=d[2] / max(n_steps, 1), value_loss=d[3], policy_loss=d[4],
                    entropy=d[5])

    def close(self):
        """Release the captured CUDA graphs.  With more than one rank the update graph holds NCCL work: drop it (and drain the
        device) BEFORE torch.distributed.destroy_process_group(), which otherwise waits forever for the communicator."""
        self._graph = None
        self._update_graph = None
        self._torch.cuda.synchronize(self.device)
        if self._peer is not None:
            self._peer.close()
            self._peer = None

    def train(self, num_updates=None, log=None):
        """The reference's outer loop: lr annealing, rollout, GAE, update; returns per-update stats."""
        torch, cfg = self._torch, self.cfg
        if num_updates is None:
            num_updates = max(1, cfg.total_timesteps // self.batch_size)
        history = []
        for update in range(1, num_updates + 1):
            self.update_idx = update - 1
            if cfg.anneal_lr:
                self._lr.fill_((1.0 - (update - 1.0) / num_updates) * cfg.learning_rate)
            t0 = time.perf_counter()
            adv, ret, ep_ret, ep_cnt = self.rollout()
            torch.cuda.synchronize(self.device)
            t1 = time.perf_counter()
            stats = self.update(adv, ret)
            torch.cuda.synchronize(self.device)
            t2 = time.perf_counter()
            y_pred, y_true = self.values.reshape(-1), ret.reshape(-1)
            var_y = y_true.var()
            stats.update(update=update, global_step=self.global_step, learning_rate=float(self._lr.item()), rollout_s=t1 - t0, update_s=t2 - t1,
                         rollout_sps=self.batch_size / (t1 - t0), mean_step_reward=self.rewards.mean().item(),
                         episodes=int(ep_cnt.item()), mean_return_per_finished_episode=(ep_ret / ep_cnt.clamp(min=1)).item(),
                         explained_variance=(1 - (y_true - y_pred).var() / var_y).item() if var_y > 0 else float("nan"))
            history.append(stats)
            if log:
                log(stats)
        return history
