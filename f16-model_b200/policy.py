"""Agent MLPs on the tensor cores -- mirror of the network part of ``Agent``
(control/ppo/ppo_model_gsde.py:37-66,81-99): ``actor_mean`` and ``critic`` are
``Linear(5,64) Tanh Linear(64,64) Tanh Linear(64,1)``; the policy is Gaussian with a
state-independent log-std.

``TensorCorePolicy`` owns ordinary torch parameters (so the PPO update can use autograd) and
evaluates them for the rollout with the hand-written tcgen05 kernel (csrc/f16_policy.cu) through
the C ABI: one launch gives mean, value, a sampled action and its log-probability for all envs.
No CPU path: the forward raises without the CUDA library / a B200."""
import math

from . import _capi


def _layer_init(layer, std=math.sqrt(2), bias_const=0.0):
    import torch

    torch.nn.init.orthogonal_(layer.weight, std)
    torch.nn.init.constant_(layer.bias, bias_const)
    return layer


_split_linear_cls = None


def split_k_linear(x, weight, bias, min_rows=8192):
    """``F.linear(x, weight, bias)`` whose backward computes the weight / bias gradients as a split-K reduction.

    The PPO update multiplies ``[64, B] x [B, 64]`` with B = 262,144 rows per minibatch: cuBLAS runs that skinny
    product as a 2-CTA SIMT GEMM (280 us, 40 % of the update) and torch's column-sum bias gradient takes 80 us.
    Splitting the batch into S slices -- ``bmm([S, out, B/S], [S, B/S, in]).sum(0)`` and ``g.view(S, B/S, out).sum(1).sum(0)``
    -- fills the GPU with S independent products in the same fp32 arithmetic (only the summation order changes)."""
    global _split_linear_cls
    import torch

    if _split_linear_cls is None:
        class _SplitKLinear(torch.autograd.Function):
            @staticmethod
            def forward(ctx, x, weight, bias, slices):
                ctx.save_for_backward(x, weight)
                ctx.slices = slices
                return torch.addmm(bias, x, weight.t())

            @staticmethod
            def backward(ctx, g):
                x, weight = ctx.saved_tensors
                S, B = ctx.slices, x.shape[0]
                g = g.contiguous()
                gs = g.view(S, B // S, g.shape[1])
                gx = g @ weight if ctx.needs_input_grad[0] else None
                gw = torch.bmm(gs.transpose(1, 2), x.view(S, B // S, x.shape[1])).sum(0)
                gb = gs.sum(1).sum(0)
                return gx, gw, gb, None

        _split_linear_cls = _SplitKLinear
    B = x.shape[0]
    if x.dim() != 2 or B < min_rows or not x.is_contiguous():
        return torch.nn.functional.linear(x, weight, bias)
    slices = 1
    while slices < 512 and B % (2 * slices) == 0 and B // (2 * slices) >= 256:
        slices *= 2
    if slices == 1:
        return torch.nn.functional.linear(x, weight, bias)
    return _split_linear_cls.apply(x, weight, bias, slices)


def mlp_forward(net, x):
    """``net(x)`` for a Sequential of Linear / Tanh layers, through ``split_k_linear``."""
    import torch

    for m in net:
        x = split_k_linear(x, m.weight, m.bias) if isinstance(m, torch.nn.Linear) else m(x)
    return x


def pack_policy_blob(actor, critic, out=None):
    """Arrange the six Linear layers of (actor, critic) in the kernel's layout (csrc/f16_policy.h,
    ``PolicyBlob``): K-major 16-byte chunks ``[chunk][row][4]`` for the two tensor-core layers."""
    import torch

    la = [m for m in actor if isinstance(m, torch.nn.Linear)]
    lc = [m for m in critic if isinstance(m, torch.nn.Linear)]
    dev = la[0].weight.device
    with torch.no_grad():
        w1 = torch.zeros(128, 8, device=dev)
        w1[:64, :5] = la[0].weight
        w1[64:, :5] = lc[0].weight
        w1p = w1.reshape(128, 2, 4).permute(1, 0, 2).contiguous()                    # [chunk][row][4]
        w2 = torch.stack([la[1].weight, lc[1].weight])                               # [2][64 out][64 in]
        w2p = w2.reshape(2, 64, 16, 4).permute(0, 2, 1, 3).contiguous()              # [net][chunk][row][4]
        b1 = torch.cat([la[0].bias, lc[0].bias])
        b2 = torch.cat([la[1].bias, lc[1].bias])
        w3 = torch.cat([la[2].weight.reshape(-1), lc[2].weight.reshape(-1)])
        b3 = torch.cat([la[2].bias.reshape(-1), lc[2].bias.reshape(-1), torch.zeros(2, device=dev)])
        flat = torch.cat([w1p.reshape(-1), w2p.reshape(-1), b1, b2, w3, b3]).float()
    n_bytes = _capi.load_library().f16_policy_blob_bytes()
    assert flat.numel() * 4 == n_bytes, (flat.numel() * 4, n_bytes)
    if out is None:
        return flat.contiguous()
    out.copy_(flat)
    return out


def pack_train_blob(actor, critic, out=None):
    """The fused PPO-gradient kernel's weights (csrc/f16_policy.h ``TrainBlob``): the forward blob followed by W2^T of
    both nets in the same K-major packing (the B operand of dH1 = dZ2 * W2)."""
    import torch

    fwd = pack_policy_blob(actor, critic)
    with torch.no_grad():
        w2t = torch.stack([[m for m in net if isinstance(m, torch.nn.Linear)][1].weight.t() for net in (actor, critic)])
        w2tp = w2t.reshape(2, 64, 16, 4).permute(0, 2, 1, 3).contiguous()            # [net][chunk over j][row i][4]
        flat = torch.cat([fwd, w2tp.reshape(-1).float()])
    assert flat.numel() * 4 == _capi.load_library().f16_ppo_train_blob_bytes()
    if out is None:
        return flat.contiguous()
    out.copy_(flat)
    return out


def pack_train_blob16(actor, critic, logical=False):
    """The fp16-operand training kernel's weights (csrc/f16_policy.h ``TrainBlob16``): W1 (K padded 5 -> 16), W2 and W2^T as
    fp16 in K-major 16-byte chunks ``[chunk][row][8 halfs]``, then the fp32 biases and output weights.  Returned as the
    fp32 words of the blob; ``logical=True`` returns (half region as float32 values, float tail) instead -- what the
    optimiser kernel's gather permutations are built from."""
    import torch

    la = [m for m in actor if isinstance(m, torch.nn.Linear)]
    lc = [m for m in critic if isinstance(m, torch.nn.Linear)]
    dev = la[0].weight.device
    with torch.no_grad():
        w1 = torch.zeros(128, 16, device=dev)
        w1[:64, :5] = la[0].weight
        w1[64:, :5] = lc[0].weight
        w1p = w1.reshape(128, 2, 8).permute(1, 0, 2)                                  # [chunk][row][8]
        w2 = torch.stack([la[1].weight, lc[1].weight])                                # [net][out j][in i]
        w2p = w2.reshape(2, 64, 8, 8).permute(0, 2, 1, 3)                             # [net][chunk over i][row j][8]
        w2tp = w2.transpose(1, 2).reshape(2, 64, 8, 8).permute(0, 2, 1, 3)            # [net][chunk over j][row i][8]
        halves = torch.cat([w1p.reshape(-1), w2p.reshape(-1), w2tp.reshape(-1)]).float()
        tail = torch.cat([la[0].bias, lc[0].bias, la[1].bias, lc[1].bias, la[2].weight.reshape(-1), lc[2].weight.reshape(-1),
                          la[2].bias.reshape(-1), lc[2].bias.reshape(-1), torch.zeros(2, device=dev)]).float()
        if logical:
            return halves, tail
        flat = torch.cat([halves.to(torch.float16).view(torch.float32), tail])
    assert flat.numel() * 4 == _capi.load_library().f16_ppo_train_blob16_bytes(), flat.numel()
    return flat.contiguous()


def scatter_table(n_params, fwd_perm, tail_perm, half_perm):
    """The inverse of the optimiser kernel's gather permutations (blob word k <- parameter perm[k] - 1, 0 = padding): for every
    parameter the word it occupies in the forward blob, in the fp32 tail of the training blob, and its (up to two: W2 and W2^T)
    fp16 words -- int32 [n_params, 4], -1 = none.  ``f16_ppo_step16`` scatters each updated parameter through it."""
    import torch

    out = torch.full((n_params, 4), -1, dtype=torch.int32, device=fwd_perm.device)
    for col, perm, mult in ((0, fwd_perm, 1), (1, tail_perm, 1), (2, half_perm, 2)):
        k = torch.nonzero(perm > 0).reshape(-1)
        j = perm[k] - 1
        order = torch.argsort(j, stable=True)
        j, k = j[order], k[order]
        first = torch.ones_like(j, dtype=torch.bool)
        first[1:] = j[1:] != j[:-1]
        start = torch.cummax(torch.where(first, torch.arange(j.numel(), device=j.device), torch.zeros_like(j)), 0).values
        rank = torch.arange(j.numel(), device=j.device) - start
        if j.numel() and int(rank.max()) >= mult:
            raise ValueError("a parameter occupies more blob words than the scatter table holds")
        out[j, col + rank] = k.to(torch.int32)
    return out.contiguous()


GRAD_FIELDS = (("w1", (2, 64, 5)), ("b1", (2, 64)), ("w2", (2, 64, 64)), ("b2", (2, 64)), ("w3", (2, 64)), ("b3", (2,)),
               ("logstd", (1,)), ("stats", (5,)))


def grad_views(flat):
    """Named views into the flat gradient buffer written by ``f16_ppo_grad`` (csrc/f16_policy.h ``PolicyGrad``)."""
    out, off = {}, 0
    for name, shape in GRAD_FIELDS:
        n = 1
        for d in shape:
            n *= d
        out[name] = flat[off:off + n].view(shape)
        off += n
    assert off == flat.numel(), (off, flat.numel())
    return out


class TensorCorePolicy:
    """actor_mean + critic + log-std, with a tensor-core forward for rollouts."""

    def __init__(self, device="cuda", seed=0, logstd_init=0.0, env_id_base=0):
        import torch
        from torch import nn

        self._torch = torch
        self._lib = _capi.load_library()
        self.device = torch.device(device if str(device) != "cuda" else f"cuda:{torch.cuda.current_device()}")
        g = torch.random.fork_rng(devices=[])
        with g:
            torch.manual_seed(seed)
            self.actor_mean = nn.Sequential(_layer_init(nn.Linear(5, 64)), nn.Tanh(), _layer_init(nn.Linear(64, 64)), nn.Tanh(),
                                            _layer_init(nn.Linear(64, 1), std=0.01)).to(self.device)
            self.critic = nn.Sequential(_layer_init(nn.Linear(5, 64)), nn.Tanh(), _layer_init(nn.Linear(64, 64)), nn.Tanh(),
                                        _layer_init(nn.Linear(64, 1), std=1.0)).to(self.device)
        self.actor_logstd = nn.Parameter(torch.full((1,), float(logstd_init), device=self.device))
        self.seed = int(seed)
        self.env_id_base = int(env_id_base)
        self._blob = pack_policy_blob(self.actor_mean, self.critic)
        self._blob16 = None      # fp16 operand blob (TrainBlob16), packed on first use; a PPOTrainer aliases its own here
        self._step = 0

    def parameters(self):
        return list(self.actor_mean.parameters()) + list(self.critic.parameters()) + [self.actor_logstd]

    def sync_weights(self):
        """Re-pack the tensor-core blob after an optimiser step."""
        pack_policy_blob(self.actor_mean, self.critic, out=self._blob)
        if self._blob16 is not None:
            self._blob16.copy_(pack_train_blob16(self.actor_mean, self.critic))

    def blob16(self):
        """The fp16 tensor-core weights (``TrainBlob16``) of ``f16_policy_forward16`` / ``f16_rollout`` / ``f16_ppo_grad16``."""
        if self._blob16 is None:
            self._blob16 = pack_train_blob16(self.actor_mean, self.critic)
        return self._blob16

    def _forward(self, precision, n):
        """(entry point, weights) of the forward kernel: "tf32" = policy_forward_kernel (one 128-row tile per 256-thread CTA),
        "fp16" = policy_forward16_kernel (the rollout kernel's schedule: four groups per CTA, biases in the MMAs), None = fp16
        from 49,152 rows up (measured, tools/policy_probe.py: 16 k rows 5.7 vs 8.0 us, 64 k rows 10.0 vs 8.1 us, 2^20 rows
        108 vs 83 us -- the 512-thread CTA has the longer prologue and the deeper pipeline)."""
        if precision is None:
            precision = "fp16" if n >= 49152 else "tf32"
        if precision == "fp16":
            return self._lib.f16_policy_forward16, self.blob16()
        if precision != "tf32":
            raise ValueError("precision must be 'tf32', 'fp16' or None")
        return self._lib.f16_policy_forward, self._blob

    # ---- tensor-core inference ------------------------------------------------------------
    def act(self, obs, sample=True, step=None, want_mean=False, out=None, env=None, precision=None):
        """obs [N,5] fp32 CUDA -> (action [N], logprob [N], value [N]) in one kernel launch.
        ``out=(action, logprob, value)`` writes into caller buffers (e.g. rows of a rollout buffer);
        ``env`` (an ``F16VecEnv``) keys the exploration noise by the env's own episode / step
        counters instead of ``step`` -- nothing host-side changes between launches, so the call can
        be captured in a CUDA graph and replayed."""
        torch = self._torch
        obs = _capi.require_cuda(obs, "obs")
        if obs.dtype != torch.float32 or obs.dim() != 2 or obs.shape[1] != 5:
            raise ValueError("obs must be float32 [N, 5]")
        n = obs.shape[0]
        if out is None:
            action = torch.empty(n, device=obs.device)
            logprob = torch.empty(n, device=obs.device)
            value = torch.empty(n, device=obs.device)
        else:
            action, logprob, value = out
        mean = torch.empty(n, device=obs.device) if want_mean else None
        if step is None and env is None:
            step = self._step
            self._step += 1
        ep_len = env.episode_length.data_ptr() if env is not None else None
        ep_no = env.episode_no.data_ptr() if env is not None else None
        fwd, blob = self._forward(precision, n)
        with torch.cuda.device(obs.device):
            _capi.check(fwd(
                n, obs.data_ptr(), blob.data_ptr(), self.actor_logstd.data_ptr(), mean.data_ptr() if want_mean else None,
                value.data_ptr(), action.data_ptr(), logprob.data_ptr(), int(bool(sample)), self.seed & 0xFFFFFFFFFFFFFFFF,
                int(step or 0) & 0xFFFFFFFF, ep_len, ep_no, self.env_id_base, _capi.stream_ptr(obs.device)))
        return (action, logprob, value, mean) if want_mean else (action, logprob, value)

    def mean_value(self, obs, precision=None):
        """(actor mean [N], critic value [N]) on the tensor cores, no sampling."""
        torch = self._torch
        obs = _capi.require_cuda(obs, "obs")
        n = obs.shape[0]
        mean = torch.empty(n, device=obs.device)
        value = torch.empty(n, device=obs.device)
        fwd, blob = self._forward(precision, n)
        with torch.cuda.device(obs.device):
            _capi.check(fwd(n, obs.data_ptr(), blob.data_ptr(), None, mean.data_ptr(), value.data_ptr(),
                            None, None, 0, 0, 0, None, None, 0, _capi.stream_ptr(obs.device)))
        return mean, value

    # ---- differentiable evaluation for the PPO update (torch autograd) ---------------------
    def evaluate(self, obs, action):
        """(logprob, entropy, value) of given actions, differentiable (ppo_model_gsde.py:81-99)."""
        torch = self._torch
        mean = mlp_forward(self.actor_mean, obs).squeeze(-1)
        logstd = self.actor_logstd.expand_as(mean)
        std = torch.exp(logstd)
        logprob = -0.5 * ((action - mean) / std) ** 2 - logstd - 0.5 * math.log(2 * math.pi)
        entropy = 0.5 + 0.5 * math.log(2 * math.pi) + logstd
        return logprob, entropy, mlp_forward(self.critic, obs).squeeze(-1)
