"""ctypes binding of the C ABI in include/f16_b200.h (the thin extension the north star
asks for: PyTorch tensors are exchanged as raw device pointers + the current CUDA stream).

The library is REQUIRED.  ``load_library()`` raises ``F16Error`` when it has not been
built; compute calls raise when there is no sm_100 device.  Nothing here falls back to
the CPU."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# F16B200_LIB: alternative build of the same library (kernel-tuning experiments only)
_LIB = os.environ.get("F16B200_LIB") or os.path.join(_HERE, "lib", "libf16b200.so")
_lib = None

F16_F32, F16_F64 = 0, 1
INIT_RANDOM, INIT_FIXED, INIT_PER_ENV = 0, 1, 2
OBS_AOS, OBS_SOA = 0, 1
FLAG_ONE_ENV_PER_THREAD, FLAG_NO_PDL, FLAG_HOST_STAGED = 1, 2, 4
ABI_VERSION = 2

EXPORTS = ["f16_abi_version", "f16_last_error", "f16_init", "f16_launch_geometry", "f16_peak_probe", "f16_env_reset", "f16_env_step",
           "f16_env_step_host", "f16_host_route", "f16_rhs", "f16_model_step", "f16_coeffs", "f16_thrust", "f16_atmosphere",
           "f16_policy_blob_bytes", "f16_policy_forward", "f16_policy_forward16", "f16_rollout", "f16_ppo_adv_stats", "f16_ppo_train_blob_bytes", "f16_ppo_grad_floats", "f16_ppo_grad", "f16_ppo_train_blob16_bytes", "f16_ppo_grad16_scratch_floats", "f16_ppo_grad16", "f16_ppo_grad16_peer", "f16_ppo_step16", "f16_ppo_step16_sync_bytes", "f16_ppo_epoch16", "f16_ppo_epoch16_sync_bytes", "f16_peer_buffer_bytes", "f16_peer_seq_words", "f16_peer_alloc", "f16_peer_open", "f16_peer_close", "f16_peer_free", "f16_ppo_pack_rows", "f16_ppo_randperm", "f16_ppo_adam",
           "f16_gae", "f16_trim"]


class F16Error(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("fast_math", C.c_int32), ("n_envs", C.c_int64), ("dt", C.c_double),
                ("horizon", C.c_int32), ("ref_len", C.c_int32), ("n_ref", C.c_int32), ("autoreset", C.c_int32),
                ("init_mode", C.c_int32), ("obs_layout", C.c_int32), ("noise", C.c_int32), ("flags", C.c_int32),
                ("seed", C.c_uint64), ("env_id_base", C.c_int64), ("x0", C.c_double * 9),
                ("seg_len", C.c_int32), ("seg_stride", C.c_int32)]


class EnvBuffers(C.Structure):
    _fields_ = [("state", C.c_void_p), ("ep_len", C.c_void_p), ("total_return", C.c_void_p), ("ref_idx", C.c_void_p),
                ("done", C.c_void_p), ("episode_no", C.c_void_p), ("ref_bank", C.c_void_p), ("init_state", C.c_void_p),
                ("pos_comp", C.c_void_p), ("seed_dev", C.c_void_p), ("ref_codes", C.c_void_p)]


class StepIO(C.Structure):
    _fields_ = [("actions", C.c_void_p), ("obs", C.c_void_p), ("reward", C.c_void_p), ("done", C.c_void_p),
                ("final_obs", C.c_void_p), ("final_ep_len", C.c_void_p), ("final_return", C.c_void_p),
                ("obs_every_step", C.c_int32), ("reserved", C.c_int32), ("stats", C.c_void_p)]


class PeerGroup(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("bufs", C.c_void_p), ("seq", C.c_void_p)]


class PpoStep16(C.Structure):
    """f16_ppo_step16_args (include/f16_b200.h): gradient kernel + fused reduction / all-reduce / clip / Adam / blob scatter."""
    _fields_ = [("n", C.c_int64), ("train_blob16", C.c_void_p), ("obs", C.c_void_p), ("actions", C.c_void_p), ("logprobs", C.c_void_p),
                ("adv", C.c_void_p), ("ret", C.c_void_p), ("val", C.c_void_p), ("mb_idx", C.c_void_p), ("logstd", C.c_void_p),
                ("adv_stats", C.c_void_p), ("grad", C.c_void_p), ("clip_coef", C.c_double), ("vf_coef", C.c_double), ("ent_coef", C.c_double),
                ("norm_adv", C.c_int32), ("clip_vloss", C.c_int32), ("scratch", C.c_void_p), ("peer", C.POINTER(PeerGroup)),
                ("n_params", C.c_int32), ("reserved", C.c_int32), ("param", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("max_exp_avg_sq", C.c_void_p), ("step", C.c_void_p), ("lr", C.c_void_p), ("beta1", C.c_double), ("beta2", C.c_double),
                ("eps", C.c_double), ("max_grad_norm", C.c_double), ("scatter", C.c_void_p), ("fwd_blob", C.c_void_p), ("tail_blob", C.c_void_p),
                ("half_blob", C.c_void_p), ("diag", C.c_void_p), ("inv_rows", C.c_double), ("sync", C.c_void_p)]


class RolloutIO(C.Structure):
    _fields_ = [("train_blob16", C.c_void_p), ("logstd", C.c_void_p), ("obs0", C.c_void_p), ("obs_next", C.c_void_p),
                ("actions", C.c_void_p), ("logprobs", C.c_void_p), ("values", C.c_void_p), ("rewards", C.c_void_p),
                ("dones", C.c_void_p), ("next_value", C.c_void_p), ("final_obs", C.c_void_p), ("final_ep_len", C.c_void_p),
                ("final_return", C.c_void_p), ("stats", C.c_void_p), ("sample", C.c_int32), ("reserved", C.c_int32),
                ("policy_seed", C.c_uint64)]


def library_path():
    return _LIB


def build(force=False, verbose=False):
    """Compile csrc/ for sm_100a with nvcc (in-tree, so the .so travels with the repo)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc")] + (["-B"] if force else [])
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode:
        print(res.stdout)
    if res.returncode:
        raise F16Error("building libf16b200.so failed (nvcc, sm_100a)")
    return _LIB


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB):
        raise F16Error(f"{_LIB} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(or `make -C f16-model_b200/csrc`). There is no CPU fallback.")
    L = C.CDLL(_LIB)
    L.f16_abi_version.restype = C.c_int
    if L.f16_abi_version() != ABI_VERSION:
        raise F16Error("libf16b200.so ABI version mismatch; rebuild")
    L.f16_last_error.restype = C.c_char_p
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    L.f16_init.argtypes = [i32]
    L.f16_launch_geometry.argtypes = [i32, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.f16_peak_probe.argtypes = [i32, i32, C.POINTER(C.c_double), vp, vp]
    L.f16_env_reset.argtypes = [C.POINTER(Config), C.POINTER(EnvBuffers), vp, i32, vp, vp]
    L.f16_env_step.argtypes = [C.POINTER(Config), C.POINTER(EnvBuffers), C.POINTER(StepIO), i32, vp]
    L.f16_env_step_host.argtypes = [C.POINTER(Config), C.POINTER(EnvBuffers), vp, vp, vp, vp, i32, i32]
    L.f16_rollout.argtypes = [C.POINTER(Config), C.POINTER(EnvBuffers), C.POINTER(RolloutIO), i32, vp]
    L.f16_rhs.argtypes = [i32, i64, vp, vp, vp, vp]
    L.f16_model_step.argtypes = [i32, i64, vp, vp, C.c_double, i32, vp]
    L.f16_coeffs.argtypes = [i32, i64, vp, vp, vp, vp, vp, vp]
    L.f16_thrust.argtypes = [i32, i64, vp, vp, vp, vp, vp]
    L.f16_atmosphere.argtypes = [i32, i64, vp, vp, vp, vp]
    L.f16_policy_blob_bytes.argtypes = []
    L.f16_ppo_adv_stats.argtypes = [i64, i32, vp, vp, vp, vp, vp]
    L.f16_ppo_train_blob_bytes.argtypes = []
    L.f16_ppo_grad_floats.argtypes = []
    L.f16_ppo_adam.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp, C.c_double, C.c_double, C.c_double, C.c_double, vp, vp, i32, vp, vp, i32,
                               vp, vp, i32, vp, vp, C.c_double, i32, vp]
    L.f16_ppo_train_blob16_bytes.argtypes = []
    L.f16_ppo_grad.argtypes = [i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_double, C.c_double, C.c_double, i32, i32, vp]
    L.f16_ppo_grad16.argtypes = L.f16_ppo_grad.argtypes[:-1] + [vp, vp]
    L.f16_ppo_grad16_scratch_floats.argtypes = [i32]
    L.f16_ppo_grad16_peer.argtypes = L.f16_ppo_grad.argtypes[:-1] + [vp, C.POINTER(PeerGroup), vp]
    L.f16_ppo_step16.argtypes = [C.POINTER(PpoStep16), vp]
    L.f16_ppo_step16_sync_bytes.argtypes = []
    L.f16_ppo_epoch16.argtypes = [C.POINTER(PpoStep16), C.c_int, vp]
    L.f16_ppo_epoch16_sync_bytes.argtypes = []
    L.f16_peer_buffer_bytes.argtypes = [i32]
    L.f16_peer_seq_words.argtypes = []
    L.f16_peer_alloc.argtypes = [i64, C.POINTER(C.c_void_p), vp]
    L.f16_peer_open.argtypes = [vp, C.POINTER(C.c_void_p)]
    L.f16_peer_close.argtypes = [vp]
    L.f16_peer_free.argtypes = [vp]
    L.f16_ppo_randperm.argtypes = [i64, C.c_uint64, vp, vp]
    L.f16_ppo_pack_rows.argtypes = [i64, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.f16_trim.argtypes = [i32, vp, vp, i32, vp, vp, vp, vp, vp]
    L.f16_gae.argtypes = [i32, i64, vp, vp, vp, vp, C.c_double, C.c_double, vp, vp, vp]
    L.f16_policy_forward.argtypes = [i64, vp, vp, vp, vp, vp, vp, vp, i32, C.c_uint64, C.c_uint32, vp, vp, i64, vp]
    L.f16_policy_forward16.argtypes = L.f16_policy_forward.argtypes
    for name in EXPORTS:
        if name not in ("f16_last_error",):
            getattr(L, name).restype = C.c_int
    L.f16_peer_buffer_bytes.restype = C.c_int64
    _lib = L
    return L


def check(rc):
    if rc != 0:
        msg = load_library().f16_last_error().decode(errors="replace")
        if rc == -1:
            raise ValueError(f"f16-model_b200: {msg}")
        raise F16Error(f"f16-model_b200 (code {rc}): {msg}")


def dtype_code(torch_dtype):
    import torch

    if torch_dtype == torch.float32:
        return F16_F32
    if torch_dtype == torch.float64:
        return F16_F64
    raise ValueError("dtype must be torch.float32 (throughput mode) or torch.float64 (parity mode)")


def require_cuda(t, name):
    if not t.is_cuda:
        raise F16Error(f"{name} must be a CUDA tensor: the F-16 simulator has no CPU path")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


def stream_ptr(device):
    import torch

    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
