"""Probe: the device-path step (f16_env_step) writing obs / reward / done straight into pinned (mapped) host memory
and reading the actions from it, against the chunked memcpy pipeline of f16_env_step_host."""
import ctypes as C, os, sys, time, torch
sys.path.insert(0, os.getcwd())
import f16_model_b200 as f16
from f16_model_b200 import _capi
CFG = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": False, "scenario": "combo"}
N = 1 << 20
for layout in ("aos", "soa"):
    for mode in ("all_mapped", "done_on_device", "reward_done_on_device", "obs_on_device"):
        env = f16.F16VecEnv(CFG, N, seed=1, fast_math=True, obs_layout=layout)
        a = (torch.rand((4, N)) * 2 - 1).pin_memory()
        a_dev = torch.empty(N, device="cuda")
        obs_h = torch.empty((N, 5) if layout == "aos" else (5, N)).pin_memory()
        rew_h = torch.empty(N).pin_memory()
        done_h = torch.empty(N, dtype=torch.uint8).pin_memory()
        obs_d = torch.empty_like(obs_h, device="cuda")
        io = _capi.StepIO()
        rew_d, done_d = torch.empty(N, device="cuda"), torch.empty(N, dtype=torch.uint8, device="cuda")
        io.reward = rew_d.data_ptr() if mode == "reward_done_on_device" else rew_h.data_ptr()
        io.done = done_d.data_ptr() if mode in ("done_on_device", "reward_done_on_device") else done_h.data_ptr()
        io.obs = obs_d.data_ptr() if mode == "obs_on_device" else obs_h.data_ptr()
        st = torch.cuda.current_stream().cuda_stream

        def step(s):
            io.actions = a[s % 4].data_ptr()
            rc = env._step_fn(env._cfg_ref, env._buf_ref, C.byref(io), 1, st)
            assert rc == 0, _capi.last_error() if hasattr(_capi, "last_error") else rc
            torch.cuda.synchronize()

        for s in range(5): step(s)
        t0 = time.perf_counter()
        for s in range(50): step(s)
        dt = time.perf_counter() - t0
        # the mapped outputs must equal what the device path writes into device buffers
        ref = f16.F16VecEnv(CFG, N, seed=1, fast_math=True, obs_layout=layout)
        for s in list(range(5)) + list(range(50)):
            o, r, d, _, _ = ref.step(a[s % 4].cuda())
        if mode != "all_mapped":
            rew_h.copy_(rew_d); done_h.copy_(done_d)
            if mode == "obs_on_device": obs_h.copy_(obs_d)
        ok = bool(torch.equal(r.cpu(), rew_h) and torch.equal(d.cpu().to(torch.uint8), done_h) and torch.equal(ref._obs.cpu(), obs_h))
        print(layout, mode, round(N * 50 / dt / 1e9, 4), "e9 env-steps/s", round(dt / 50 * 1e3, 4), "ms/step", "equal" if ok else "MISMATCH", flush=True)
