"""Env layer -- vectorised drop-in for ``F16model.env.F16`` (F16model/env/env.py:13-262).

``F16VecEnv`` plays the role of ``gym.vector.SyncVectorEnv([F16(config)] * N)`` as the PPO
trainer uses it (control/ppo/ppo_model_gsde.py:201-230), but all N aircraft advance in ONE
CUDA kernel launch and observations, rewards and termination flags stay on the device:

    envs = F16VecEnv(config, num_envs=1 << 20, device="cuda:0")
    obs, _ = envs.reset(seed=1)                     # obs: cuda tensor [N, 5]
    obs, reward, terminated, truncated, info = envs.step(actions)   # actions: cuda tensor [N] / [N, 1]

``F16`` is the single-env facade with the reference's NumPy in / NumPy out signature.
There is no CPU fallback: without the CUDA library or a B200 the constructors raise."""
import ctypes as C

import numpy as np

from . import _capi, plane
from .model import States, Control
from .task import ReferenceSignal, build_combo_segments, build_reference_bank, signals_from_codes


def find_correct_thrust_position(control_throttle):
    """Engine power [%] at which ``power_level(., throttle)`` vanishes (model/engine.py:35-38).
    The reference finds it with fsolve; the root is the commanded power level in closed form."""
    t = float(control_throttle)
    return 64.94 * t if t <= 0.77 else 217.38 * t - 117.38


def get_trimmed_state_control():
    """Trim for V = 200 m/s, H = 3000 m (env/env.py:246-262): returns (x0[9], u0[2])."""
    u = Control(np.radians(-4.3674), 0.3767)
    x0 = States(Ox=0, Oy=3000, wz=0, theta=np.radians(2.7970), V=200, alpha=np.radians(2.7970), stab=u.stab,
                dstab=np.radians(0), Pa=find_correct_thrust_position(u.throttle))
    return x0.to_array(), u.to_array()


def get_action():
    """A random control (env/env.py:217-218): stabilator U(-25, 25) deg in radians, throttle U(0, 1),
    drawn from the process-global ``np.random`` in the reference's order."""
    return Control(np.radians(np.random.uniform(-25, 25)), np.random.uniform(0, 1))


def get_random_state():
    """One draw of the reference's random initial state on the HOST with ``np.random``
    (env/env.py:221-243), same draw order: alpha(=theta), Oy, wz, V."""
    b = plane.random_state_bound
    alpha = np.random.uniform(b["alpha"][0], b["alpha"][1], 1)[0]
    Oy = np.random.uniform(b["Oy"][0], b["Oy"][1], 1)[0]
    wz = np.random.uniform(b["wz"][0], b["wz"][1], 1)[0]
    V = np.random.uniform(b["V"][0], b["V"][1], 1)[0]
    return States(Ox=0, Oy=Oy, wz=wz, theta=alpha, V=V, alpha=alpha, stab=0, dstab=0,
                  Pa=find_correct_thrust_position(0))


class Box:
    """The two attributes of ``gymnasium.spaces.Box`` the trainer reads."""

    def __init__(self, low, high, shape, dtype=np.float32):
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.low = np.full(self.shape, low, dtype=dtype)
        self.high = np.full(self.shape, high, dtype=dtype)


def _normalize(truncated_state):
    return [(truncated_state[i] - lo) / (hi - lo) for i, (lo, hi) in enumerate(plane.state_bound.values())]


def _denormalize(state_norm):
    vals = [state_norm[i] * (hi - lo) + lo for i, (lo, hi) in enumerate(plane.state_bound.values())]
    out = np.array(vals)
    k = len(vals)
    if k < len(state_norm):
        out = np.concatenate((out, state_norm[k:]))
    return out


def _rescale_action(action):
    """[-1, 1] stick -> stabilator command [rad] (env/env.py:200-211, utils/calc.py:4-11):
    the bounds are cast to float32 and subtracted in float32, exactly like the reference."""
    a = np.clip(action[0], -1, 1)
    low, high = np.float32(-plane.maxabsstab), np.float32(plane.maxabsstab)
    return low + 0.5 * (a + 1.0) * (high - low)


class F16VecEnv:
    """N F-16 longitudinal envs advanced by one fused CUDA kernel per call.

    Parameters
    ----------
    config : dict -- the reference's keys: ``dt``, ``tn``, ``debug_state``, ``determenistic_ref``,
        ``scenario`` (None | "step" | "cos" | "pure_step" | "combo"), optional ``init_state``
        (ndarray[6]) + ``init_control`` (ndarray[2]) or, with ``debug_state``, a ``States``;
        optional ``noise``.  Extension: ``init_state`` may be a [9, N] / [N, 9] tensor or array
        of explicit per-env states.  Without ``init_state`` each episode starts from
        ``get_random_state()`` drawn on the device.
    num_envs, device, dtype : batch size, CUDA device, torch.float32 (throughput) or
        torch.float64 (parity mode, 1e-10 vs the reference).
    autoreset : gymnasium ``SyncVectorEnv`` semantics (reset inside the step, return the reset
        observation, terminal observation in ``info["final_observation"]``).  With False a
        finished env freezes until ``reset``.
    fast_math : fp32 only -- MUFU intrinsics for sin/cos/pow/div.
    compensated_position : keep a Kahan carry for the Ox / Oy position integrals (default off; +16 B per
        env-step).  Plain fp32 accumulation of ~2 m increments onto ~2000 m rounds once per step; the carry
        removes that drift (median Oy error 7e-7 -> 6e-8 over an episode).  Meant for fp32 runs much
        longer than the reference's 1000-step episodes.
    env_id_base : global id of this shard's env 0 (multi-GPU: rank * num_envs); the device RNG is
        keyed by global env id, so a sharded run reproduces the single-GPU run env for env.
    obs_layout : "aos" -> obs [N, 5] contiguous (reference layout); "soa" -> stored [5, N],
        returned as the transposed view [N, 5].
    bank_size : rows of the device reference bank for random scenarios.  None (default) = the reference's own
        distribution: random ``combo`` episodes are served from the 52 KB segment bank that holds all 6480 signals
        (``ref_mode == "segments"``: ``ref_idx`` then carries packed segment codes, ``reference_signals()`` rebuilds
        rows); the other random scenarios are enumerated in full.  An integer (or "full") asks for a bank of
        whole rows instead, see ``task.build_reference_bank``.
    two_per_thread, pdl, host_zero_copy : per-env kernel / route selection (``f16_config.flags``): the
        one-env-per-thread step kernel instead of the two-envs-per-thread one, launches without programmatic
        dependent launch, the staged host route instead of zero-copy.  All variants are bit-identical
        (tests/test_gpu_parity.py); the switches exist for tests and profiling.
    """

    metadata = {}

    def __init__(self, config, num_envs, device="cuda", dtype=None, autoreset=True, seed=0, fast_math=False,
                 obs_layout="aos", bank_size=None, reference_bank=None, env_id_base=0, compensated_position=None,
                 two_per_thread=True, pdl=True, host_zero_copy=True):
        import torch

        self._torch = torch
        self._lib = _capi.load_library()
        if not torch.cuda.is_available():
            raise _capi.F16Error("no CUDA device: f16-model_b200 has no CPU fallback")
        self.config = config
        self.num_envs = int(num_envs)
        if self.num_envs <= 0:
            raise ValueError("num_envs must be positive")
        self.device = torch.device(device if str(device) != "cuda" else f"cuda:{torch.cuda.current_device()}")
        self.dtype = dtype or torch.float32
        self.dt = float(config["dt"])
        self.tn = float(config["tn"])
        self.autoreset = bool(autoreset)
        self.obs_layout = obs_layout
        N, dev, dt_ = self.num_envs, self.device, self.dtype
        _capi.check(self._lib.f16_init(self.device.index))

        # ---- reference signals (env/task.py) --------------------------------
        # Random `combo` episodes (the reference's training scenario) are served from the SEGMENT bank: all 6480 signals of
        # the reference's distribution in 52 KB (task.build_combo_segments).  Everything else -- deterministic scenarios, the
        # small enumerable scenarios, a caller-supplied bank, an explicit bank_size -- is a bank of whole rows [R, T].
        segments = None
        if reference_bank is None and bank_size is None and config["scenario"] == "combo" and not config["determenistic_ref"]:
            segments = build_combo_segments(self.dt, self.tn)
        self.ref_codes = None
        if segments is not None:
            seg_bank, codes, seg_len = segments
            ref_len = int(np.arange(0, self.tn, self.dt).size)
            self.ref_mode = "segments"
            self._seg_bank_host, self._codes_host = seg_bank, codes
            self._seg_stride = (seg_bank.shape[0] + 7) // 8 * 8            # rows of one sample padded to whole 32-byte sectors
            padded = np.zeros((seg_len, self._seg_stride))
            padded[:, : seg_bank.shape[0]] = seg_bank.T                     # sample-major: the envs of a warp share the sample index
            self.ref_bank = torch.tensor(padded, dtype=dt_, device=dev)
            self.ref_codes = torch.tensor(codes.astype(np.int64), dtype=torch.int64, device=dev).to(torch.int32)
            n_ref = int(codes.size)
        else:
            if reference_bank is None:
                reference_bank = build_reference_bank(self.dt, self.tn, config["scenario"], config["determenistic_ref"],
                                                      bank_size=bank_size, seed=seed)
            bank = np.ascontiguousarray(np.atleast_2d(np.asarray(reference_bank, dtype=np.float64)))
            self.ref_mode = "rows"
            self.ref_bank_host = bank
            self.ref_bank = torch.tensor(bank, dtype=dt_, device=dev)
            ref_len, n_ref, seg_len = bank.shape[1], bank.shape[0], 0
            self._seg_stride = 0
        horizon = self.tn / self.dt - 1          # env.py:115 compares the step counter with this float
        self.horizon = int(horizon) if float(horizon).is_integer() else ref_len - 1

        # ---- initial state policy (env.py:172-195) ---------------------------
        self._x0 = np.zeros(9)
        init_mode = _capi.INIT_RANDOM
        self.init_state_tensor = None
        if config.get("debug_state"):
            s = config["init_state"]
            self._x0 = np.asarray(s.to_array(), dtype=np.float64).reshape(9)
            init_mode = _capi.INIT_FIXED
        elif "init_state" in config:
            s = config["init_state"]
            if isinstance(s, np.ndarray) and s.ndim == 1 and isinstance(config.get("init_control"), np.ndarray):
                u = config["init_control"]
                self._x0 = np.array([s[0], s[1], s[2], s[3], s[4], s[5], u[0], 0.0,
                                     find_correct_thrust_position(u[1])], dtype=np.float64)
                init_mode = _capi.INIT_FIXED
            else:
                t = torch.as_tensor(s)
                if t.dim() == 2 and t.shape == (N, 9) and N != 9:
                    t = t.t()
                if t.dim() != 2 or t.shape != (9, N):
                    raise ValueError(f"per-env init_state must have shape [9, {N}] (or [{N}, 9])")
                self.init_state_tensor = t.to(device=dev, dtype=dt_).contiguous()
                init_mode = _capi.INIT_PER_ENV

        # ---- persistent device buffers ---------------------------------------
        self.state = torch.zeros((9, N), dtype=dt_, device=dev)
        self.episode_length = torch.zeros(N, dtype=torch.int32, device=dev)
        self.total_return = torch.zeros(N, dtype=dt_, device=dev)
        self.ref_idx = torch.zeros(N, dtype=torch.int32, device=dev)
        self.done = torch.zeros(N, dtype=torch.uint8, device=dev)
        self.episode_no = torch.zeros(N, dtype=torch.int32, device=dev)
        # the RNG key lives on the device too: kernels captured in a CUDA graph see a later reset(seed=...)
        self._seed_dev = torch.full((1,), (int(seed) & 0xFFFFFFFFFFFFFFFF) - (1 << 64 if int(seed) & (1 << 63) else 0),
                                    dtype=torch.int64, device=dev)
        compensated_position = bool(compensated_position) and dt_ == torch.float32
        self.pos_comp = torch.zeros((2, N), dtype=dt_, device=dev) if compensated_position else None
        obs_shape = (N, 5) if obs_layout == "aos" else (5, N)
        self._obs = torch.zeros(obs_shape, dtype=dt_, device=dev)
        self._final_obs = torch.zeros(obs_shape, dtype=dt_, device=dev)
        self._final_len = torch.zeros(N, dtype=torch.int32, device=dev)
        self._final_ret = torch.zeros(N, dtype=dt_, device=dev)
        self._reward = torch.zeros((1, N), dtype=dt_, device=dev)
        self._done_out = torch.zeros((1, N), dtype=torch.uint8, device=dev)
        self._truncated = torch.zeros(N, dtype=torch.bool, device=dev)
        # in-kernel episode statistics (warp ballot + shuffle reduction, one atomic set per warp with a termination)
        self.episode_stats = torch.zeros(4, dtype=torch.float64, device=dev)

        cfg = _capi.Config()
        cfg.dtype = _capi.dtype_code(dt_)
        cfg.fast_math = int(bool(fast_math) and dt_ == torch.float32)
        cfg.n_envs = N
        cfg.dt = self.dt
        cfg.horizon = self.horizon
        cfg.ref_len = ref_len
        cfg.n_ref = n_ref
        cfg.seg_len = seg_len
        cfg.seg_stride = self._seg_stride
        cfg.autoreset = int(self.autoreset)
        cfg.init_mode = init_mode
        cfg.obs_layout = _capi.OBS_AOS if obs_layout == "aos" else _capi.OBS_SOA
        cfg.noise = int(bool(config.get("noise", False)))
        cfg.flags = ((0 if two_per_thread else _capi.FLAG_ONE_ENV_PER_THREAD) | (0 if pdl else _capi.FLAG_NO_PDL)
                     | (0 if host_zero_copy else _capi.FLAG_HOST_STAGED))
        cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        cfg.env_id_base = int(env_id_base)
        for j in range(9):
            cfg.x0[j] = float(self._x0[j])
        self._cfg = cfg
        self._bind()

        # gym-style spaces (env.py:32-35): action Box(-1,1,(1,),f32), observation Box(-1,1,(5,))
        self.single_action_space = Box(-1, 1, (1,), np.float32)
        self.single_observation_space = Box(-1, 1, (5,), np.float32)
        self.action_space = Box(-1, 1, (N, 1), np.float32)
        self.observation_space = Box(-1, 1, (N, 5), np.float32)
        self.reset(seed=seed)

    # ------------------------------------------------------------------ plumbing
    def _bind(self):
        b = _capi.EnvBuffers()
        b.state = self.state.data_ptr()
        b.ep_len = self.episode_length.data_ptr()
        b.total_return = self.total_return.data_ptr()
        b.ref_idx = self.ref_idx.data_ptr()
        b.done = self.done.data_ptr()
        b.episode_no = self.episode_no.data_ptr()
        b.ref_bank = self.ref_bank.data_ptr()
        b.init_state = self.init_state_tensor.data_ptr() if self.init_state_tensor is not None else None
        b.pos_comp = self.pos_comp.data_ptr() if self.pos_comp is not None else None
        b.seed_dev = self._seed_dev.data_ptr()
        b.ref_codes = self.ref_codes.data_ptr() if self.ref_codes is not None else None
        self._buf = b
        io = _capi.StepIO()
        io.obs = self._obs.data_ptr()
        io.reward = self._reward.data_ptr()
        io.done = self._done_out.data_ptr()
        io.final_obs = self._final_obs.data_ptr()
        io.final_ep_len = self._final_len.data_ptr()
        io.final_return = self._final_ret.data_ptr()
        io.obs_every_step = 0
        io.stats = self.episode_stats.data_ptr()
        self._io1 = io
        # prebuilt call arguments / results of the K = 1 fast path
        self._Tensor = self._torch.Tensor
        self._step_fn = self._lib.f16_env_step
        self._cfg_ref, self._buf_ref, self._io1_ref = C.byref(self._cfg), C.byref(self._buf), C.byref(self._io1)
        terminated = self._done_out[0].view(self._torch.bool)
        info = {
            "episode_length": self.episode_length,
            "total_return": self.total_return,
            "state": self.state,
            "_final_info": terminated,
            "_final_observation": terminated,
            "final_observation": self._obs_view(self._final_obs),
            "final_info": {"episode_length": self._final_len, "total_return": self._final_ret},
        }
        self._step_result = (self._obs_view(self._obs), self._reward[0], terminated, self._truncated, info)

    def _stream(self):
        return C.c_void_p(self._torch.cuda.current_stream(self.device).cuda_stream)

    def _obs_view(self, obs):
        return obs if self.obs_layout == "aos" else obs.transpose(-1, -2)

    def _actions(self, actions, K=None):
        torch = self._torch
        if isinstance(actions, np.ndarray):
            actions = torch.as_tensor(actions).to(self.device, non_blocking=True)
        if not isinstance(actions, torch.Tensor):
            raise TypeError(f"Action has type '{type(actions)}' should be torch.Tensor or np.ndarray")
        if not actions.is_cuda:
            actions = actions.to(self.device, non_blocking=True)
        if actions.dtype != self.dtype:
            actions = actions.to(self.dtype)
        N = self.num_envs
        if K is None:
            if actions.numel() != N:
                raise ValueError(f"expected {N} actions, got shape {tuple(actions.shape)}")
        elif actions.numel() != K * N or actions.shape[0] != K:
            raise ValueError(f"expected actions of shape [{K}, {N}], got {tuple(actions.shape)}")
        return actions.contiguous()

    # ------------------------------------------------------------------ gym API
    def reset(self, seed=None, options=None, mask=None, keep_reference=False, keep_state=False):
        """``F16.reset`` for every env (env.py:154-170), or for those where ``mask`` is set.
        ``seed`` re-keys the device RNG (the reference reseeds the process-global RNGs).
        ``keep_state`` keeps the states the caller has just loaded (``set_state``) and only clears the
        counters and computes the reset observation on the device."""
        torch = self._torch
        if seed is not None:
            key = int(seed) & 0xFFFFFFFFFFFFFFFF
            self._cfg.seed = key
            self._seed_dev.fill_(key - (1 << 64) if key & (1 << 63) else key)   # stream-ordered: later graph replays see it
        mptr = None
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
            mptr = mask.data_ptr()
        with torch.cuda.device(self.device):
            _capi.check(self._lib.f16_env_reset(C.byref(self._cfg), C.byref(self._buf), mptr, int(bool(keep_reference)) | (2 if keep_state else 0),
                                                self._obs.data_ptr(), self._stream()))
        if mask is None:
            self.init_state = self.state.clone()
        return self._obs_view(self._obs), {}

    def step(self, actions):
        """One ``F16.step`` for every env (env.py:46-70) -> (obs [N,5], reward [N], terminated [N] bool,
        truncated [N] bool (all False, as in the reference), info).

        The returned tensors (and those in ``info``) are the env's own output buffers: the next
        ``step`` overwrites them, so copy what you keep (a rollout buffer does)."""
        a = actions
        if not (a.__class__ is self._Tensor and a.dtype == self.dtype and a.device == self.device
                and a.is_contiguous() and a.numel() == self.num_envs):
            a = self._actions(a)
        self._io1.actions = a.data_ptr()
        stream = self._torch.cuda.current_stream(self.device).cuda_stream
        if self._torch.cuda.current_device() == self.device.index:
            rc = self._step_fn(self._cfg_ref, self._buf_ref, self._io1_ref, 1, stream)
        else:
            with self._torch.cuda.device(self.device):
                rc = self._step_fn(self._cfg_ref, self._buf_ref, self._io1_ref, 1, stream)
        if rc:
            _capi.check(rc)
        return self._step_result

    def make_graphed_step(self, actions):
        """Capture ``step(actions)`` into a CUDA graph (launch-bound inner loops: one graph launch
        instead of a Python call + kernel launch).  ``actions`` is a static device buffer the
        caller refills before each ``replay()``; outputs land in the usual step buffers."""
        torch = self._torch
        a = self._actions(actions)
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            self.step(a)                    # warm up outside capture (table upload, lazy init)
        torch.cuda.current_stream(self.device).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.step(a)
        return graph, self._step_result

    def step_k(self, actions, obs_every_step=False, out=None):
        """K consecutive steps in ONE launch with the state held in registers.  actions [K, N]
        -> (obs, reward [K, N], done [K, N] uint8); obs is the last observation [N, 5], or
        [K, N, 5] with ``obs_every_step``.  ``out`` may carry preallocated (obs, reward, done)."""
        torch = self._torch
        if getattr(actions, "ndim", 0) != 2:
            raise ValueError("step_k takes actions of shape [K, N]")
        K = int(actions.shape[0])
        a = self._actions(actions, K)
        N, dev, dt_ = self.num_envs, self.device, self.dtype
        if out is None:
            if obs_every_step:
                obs = torch.empty((K, N, 5) if self.obs_layout == "aos" else (K, 5, N), dtype=dt_, device=dev)
            else:
                obs = self._obs
            reward = torch.empty((K, N), dtype=dt_, device=dev)
            done = torch.empty((K, N), dtype=torch.uint8, device=dev)
        else:
            obs, reward, done = out
        io = _capi.StepIO()
        io.actions = a.data_ptr()
        io.obs = obs.data_ptr()
        io.reward = reward.data_ptr()
        io.done = done.data_ptr()
        io.final_obs = self._final_obs.data_ptr()
        io.final_ep_len = self._final_len.data_ptr()
        io.final_return = self._final_ret.data_ptr()
        io.obs_every_step = int(obs_every_step)
        io.stats = self.episode_stats.data_ptr()
        with torch.cuda.device(dev):
            _capi.check(self._lib.f16_env_step(C.byref(self._cfg), C.byref(self._buf), C.byref(io), K, self._stream()))
        return self._obs_view(obs), reward, done

    def step_host(self, actions, out=None):
        """End-to-end step through HOST memory: ``actions`` is a NumPy array / CPU tensor [N] or
        [K, N]; returns (obs [N,5], reward [K,N], done [K,N]) as CPU tensors (``out`` may carry preallocated
        ones).  With pinned actions and outputs the fp32 gym step takes the library's zero-copy route (the
        kernel reads / bulk-stores host memory itself); otherwise the copies are chunked and overlapped
        with the kernel (f16_env_step_host).  ``last_host_route`` says which: "zero_copy" or "staged"."""
        torch = self._torch
        a = torch.as_tensor(actions)
        if a.is_cuda:
            raise ValueError("step_host takes host memory; use step()/step_k() for device tensors")
        a = a.to(self.dtype).contiguous()
        K = 1 if a.numel() == self.num_envs else int(a.shape[0])
        if a.numel() != K * self.num_envs:
            raise ValueError("bad action shape")
        if out is None:
            obs = torch.empty(tuple(self._obs.shape), dtype=self.dtype).pin_memory()
            reward = torch.empty((K, self.num_envs), dtype=self.dtype).pin_memory()
            done = torch.empty((K, self.num_envs), dtype=torch.uint8).pin_memory()
        else:
            obs, reward, done = out
        torch.cuda.current_stream(self.device).synchronize()   # the library pipelines on its own streams
        _capi.check(self._lib.f16_env_step_host(C.byref(self._cfg), C.byref(self._buf), a.data_ptr(), obs.data_ptr(),
                                                reward.data_ptr(), done.data_ptr(), K, self.device.index))
        self.last_host_route = "zero_copy" if self._lib.f16_host_route() == 1 else "staged"
        return self._obs_view(obs), reward, done

    # ------------------------------------------------------------------ state access
    def set_state(self, state, episode_length=None, ref_idx=None):
        """Load explicit states [9, N] (or [N, 9]) and clear the done flags / returns."""
        torch = self._torch
        t = torch.as_tensor(state)
        N = self.num_envs
        if t.dim() != 2 or (tuple(t.shape) != (9, N) and tuple(t.shape) != (N, 9)):
            raise ValueError(f"state must have shape [9, {N}] (or [{N}, 9]), got {tuple(t.shape)}")
        if tuple(t.shape) != (9, N):     # N == 9 is read as [9, N] = field-major, the storage layout
            t = t.t()
        self.state.copy_(t.to(device=self.device, dtype=self.dtype))
        self.done.zero_()
        self.total_return.zero_()
        if self.pos_comp is not None:
            self.pos_comp.zero_()
        if episode_length is None:
            self.episode_length.zero_()
        else:
            self.episode_length.copy_(torch.as_tensor(episode_length, dtype=torch.int32))
        if ref_idx is not None:
            self.ref_idx.copy_(torch.as_tensor(ref_idx, dtype=torch.int32))
        self.init_state = self.state.clone()

    def reference_signals(self, ref_idx=None):
        """``wz_ref`` rows [n, T] (NumPy, fp64) followed by the given envs' ``ref_idx`` values (default: every env's current
        one) -- the rows of the bank, or, in segment mode, the signals rebuilt from the packed segment codes."""
        idx = self.ref_idx.cpu().numpy() if ref_idx is None else np.atleast_1d(np.asarray(ref_idx))
        if self.ref_mode == "segments":
            return signals_from_codes(self._seg_bank_host, idx.astype(np.int64).astype(np.uint32), self._cfg.ref_len)
        return self.ref_bank_host[idx]

    def pop_episode_stats(self):
        """(episodes finished, mean return, mean length) since the last call -- the counters the kernel keeps
        with warp-level reductions (replaces the per-env Python loop of control/ppo/utils.py:197-235)."""
        s = self.episode_stats.clone()
        self.episode_stats.zero_()
        n = s[0].clamp(min=1.0)
        return int(s[0].item()), (s[1] / n).item(), (s[2] / n).item()

    @property
    def clock(self):
        """Simulated time per env, ``round(episode_length * dt, 4)`` (env.py:54-55)."""
        return self._torch.round(self.episode_length.to(self._torch.float64) * self.dt, decimals=4)

    @property
    def prev_state(self):
        return self.state

    def call(self, name):
        """``SyncVectorEnv.call``: tuple of per-env attributes (small batches only)."""
        if self.num_envs > 65536:
            raise ValueError("call() materialises one Python object per env; read the batched tensors instead")
        if name == "init_state":
            arr = self.init_state.double().cpu().numpy()
            return tuple(States.from_array(arr[:, i]) for i in range(self.num_envs))
        if name == "ref_signal":
            return tuple(_BankSignal(row, self.dt, self.tn) for row in self.reference_signals())
        if name == "episode_length":
            return tuple(self.episode_length.cpu().tolist())
        if name == "total_return":
            return tuple(self.total_return.cpu().tolist())
        raise AttributeError(name)

    def close(self):
        pass

    normalize = staticmethod(_normalize)
    denormalize = staticmethod(_denormalize)
    rescale_action = staticmethod(_rescale_action)


class _BankSignal:
    """Row of the reference bank presented with ``ReferenceSignal``'s attributes."""

    def __init__(self, wz_ref, dt, tn):
        self.dt, self.tn = dt, tn
        self.t = np.arange(0, tn, dt)
        self.wz_ref = wz_ref
        self.theta_ref = wz_ref * 2


class F16:
    """Single-env facade with the reference's exact call signature (env/env.py:13-70):
    ``step(np.ndarray[1]) -> (obs ndarray[5], reward float, done bool, False, info)``.
    Runs the same CUDA kernel with N = 1 in fp64 (the parity mode).

    One deliberate difference: after ``done`` the reference keeps integrating if ``step`` is called again (every
    caller in the reference resets instead); here a finished episode is frozen -- reward 0, state unchanged,
    ``done`` stays True -- until ``reset``."""

    def __init__(self, config, device="cuda", dtype=None, seed=0):
        import torch

        self._torch = torch
        self.config = config
        self.dt = float(config["dt"])
        self.tn = float(config["tn"])
        # one signal per episode, drawn on the host exactly like the reference does
        self.ref_signal = ReferenceSignal(self.dt, self.tn, determenistic=config["determenistic_ref"],
                                          scenario=config["scenario"])
        ref = self.ref_signal.wz_ref[None]
        self._vec = F16VecEnv(config, 1, device=device, dtype=dtype or torch.float64, autoreset=False, seed=seed,
                              reference_bank=ref)
        self.action_space = Box(-1, 1, (1,), np.float32)
        self.observation_space = Box(-1, 1, (5,), np.float64)
        if self._vec._cfg.init_mode == _capi.INIT_RANDOM:
            # like the reference's constructor (env/env.py:27 -> :172-195), the first initial state comes from the
            # process-global np.random, so np.random.seed(s); F16(cfg) reproduces the reference's first episode
            self._vec.set_state(torch.tensor(get_random_state().to_array().reshape(9, 1)))
            self._vec.reset(keep_state=True, keep_reference=True)
        self._sync_init()

    def _sync_init(self):
        self.init_state = States.from_array(self._vec.state[:, 0].double().cpu().numpy())
        self.prev_state = self.init_state
        self.clock = 0
        self.total_return = 0
        self.episode_length = 0
        self.done = False

    def reset(self, seed=None, options=None):
        import random

        if seed:
            random.seed(seed)
            np.random.seed(seed)
        self.ref_signal = ReferenceSignal(self.dt, self.tn, determenistic=self.config["determenistic_ref"],
                                          scenario=self.config["scenario"])
        v = self._vec
        v.ref_bank.copy_(self._torch.tensor(self.ref_signal.wz_ref[None], dtype=v.dtype))
        v.ref_bank_host = self.ref_signal.wz_ref[None]
        if v._cfg.init_mode == _capi.INIT_RANDOM:
            # reference semantics: the initial state comes from the (re)seeded global np.random
            v.set_state(self._torch.tensor(get_random_state().to_array().reshape(9, 1)))
            obs, _ = v.reset(keep_state=True, keep_reference=True)
        else:
            obs, _ = v.reset(seed=seed)
        obs = obs[0].double().cpu().numpy()
        self._sync_init()
        return obs, {}

    def step(self, action):
        if not isinstance(action, np.ndarray):
            raise TypeError(f"Action has type '{type(action)}' should be np.array")
        if self.episode_length >= self._vec._cfg.ref_len:
            raise IndexError("episode stepped past the end of the reference signal")
        a = self._torch.tensor(np.asarray(action, dtype=np.float64).reshape(1)[:1])
        obs, reward, done, _, info = self._vec.step(a)
        self.prev_state = States.from_array(self._vec.state[:, 0].double().cpu().numpy())
        self.episode_length = int(self._vec.episode_length[0])
        self.total_return = float(self._vec.total_return[0])
        self.clock = round(self.episode_length * self.dt, 4)
        self.done = bool(done[0])
        info = {"episode_length": self.episode_length, "total_return": self.total_return, "clock": self.clock,
                "state": self.prev_state}
        return obs[0].double().cpu().numpy(), float(reward[0]), self.done, False, info

    def render(self):
        raise UserWarning("TODO: Implement this function")

    def state_transform(self, state):
        """Observation of a ``States`` at the current step (env/env.py:83-101): normalised [Oy, wz, V], the
        reference sample ``wz_ref[episode_length]`` and the tracking error ``0.5 * (wz_ref - wz)``.
        Host-side helper (NumPy, fp64); ``step`` gets the same five numbers from the kernel."""
        short = _normalize([vars(state)[k] for k in plane.state_bound if k in vars(state)])
        ref = self.ref_signal.wz_ref[self.episode_length]
        short.append(ref)
        short.append(0.5 * (ref - state.wz))
        return np.array(short)

    def _state_size(self):
        """env/env.py:212-213."""
        return self.state_transform(self.init_state).size

    normalize = staticmethod(_normalize)
    denormalize = staticmethod(_denormalize)
    rescale_action = staticmethod(_rescale_action)
