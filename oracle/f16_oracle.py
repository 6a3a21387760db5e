"""ctypes front-end of the C oracle (oracle/f16_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of f16_oracle.c for the reference
file:line map and the parity-pinning status.  Only tests/,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libf16_oracle.so")
_lib = None

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile the oracle shared library (gcc, no FMA contraction)."""
    src = os.path.join(_HERE, "f16_oracle.c")
    hdr = os.path.join(_HERE, "f16_oracle_tables.h")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "libf16_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_isa.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_isa.restype = C.c_int
        L.orc_thrust.argtypes = [C.c_double] * 3
        L.orc_thrust.restype = C.c_double
        L.orc_power_level.argtypes = [C.c_double] * 2
        L.orc_power_level.restype = C.c_double
        L.orc_thrust_position.argtypes = [C.c_double]
        L.orc_thrust_position.restype = C.c_double
        L.orc_rescale_action.argtypes = [C.c_double]
        L.orc_rescale_action.restype = C.c_double
        L.orc_reward.argtypes = [C.c_double] * 2
        L.orc_reward.restype = C.c_double
        L.orc_get_pchip.argtypes = [C.c_int, _dp]
        L.orc_coeffs_batch.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _dp]
        L.orc_thrust_batch.argtypes = [C.c_int, _dp, _dp, _dp, _dp]
        L.orc_isa_batch.argtypes = [C.c_int, _dp, _dp, _dp]
        L.orc_rhs_batch.argtypes = [C.c_int, _dp, _dp, _dp]
        L.orc_rollout.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int, _dp, _dp, _dp, _ip, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _ip, _dp, _dp, C.c_int]
        L.orc_rollout.restype = C.c_longlong
        L.orc_trim_cost.argtypes = [C.c_double, C.c_double, _dp]
        L.orc_trim_cost.restype = C.c_double
        L.orc_trim.argtypes = [C.c_int, _dp, _dp, C.c_int, _dp, _dp, _dp, _ip, C.c_int]
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def isa(h):
    h = np.atleast_1d(_f64(h))
    rho, a = np.empty_like(h), np.empty_like(h)
    lib().orc_isa_batch(h.size, h, rho, a)
    return rho, a


def coeffs(alpha, fi, wz, V):
    alpha, fi, wz, V = (np.atleast_1d(_f64(v)) for v in (alpha, fi, wz, V))
    out = np.empty((alpha.size, 3))
    lib().orc_coeffs_batch(alpha.size, alpha, fi, wz, V, out)
    return out


def thrust(H, M, Pa):
    H, M, Pa = (np.atleast_1d(_f64(v)) for v in (H, M, Pa))
    out = np.empty_like(H)
    lib().orc_thrust_batch(H.size, H, M, Pa, out)
    return out


def power_level(Pa, thr):
    return np.array([lib().orc_power_level(float(p), float(t)) for p, t in zip(np.atleast_1d(Pa), np.atleast_1d(thr))])


def thrust_position(thr):
    return lib().orc_thrust_position(float(thr))


def rescale_action(a):
    return np.array([lib().orc_rescale_action(float(v)) for v in np.atleast_1d(a)])


def reward(wz_ref, wz):
    return np.array([lib().orc_reward(float(r), float(w)) for r, w in zip(np.atleast_1d(wz_ref), np.atleast_1d(wz))])


def pchip_coeffs(which):
    n = 16 if which == 3 else 76
    out = np.empty(n)
    lib().orc_get_pchip(which, out)
    return out.reshape(4, -1)


def rhs(x, u):
    x = _f64(x).reshape(-1, 9)
    u = _f64(u).reshape(-1, 2)
    dx = np.empty_like(x)
    lib().orc_rhs_batch(x.shape[0], x, u, dx)
    return dx


def rollout(x0, actions, ref_bank, ref_idx, dt=0.01, horizon=999, want_states=True, want_obs=True,
            nthreads=0):
    """Roll envs (frozen after done). x0 [n,9]; actions [T,n]; ref_bank [R,Tref];
    ref_idx [n]. Returns dict of arrays (see orc_rollout)."""
    x0 = _f64(x0).reshape(-1, 9)
    n = x0.shape[0]
    actions = _f64(actions).reshape(-1, n)
    T = actions.shape[0]
    ref_bank = _f64(ref_bank)
    if ref_bank.ndim == 1:
        ref_bank = ref_bank[None]
    ref_idx = np.ascontiguousarray(np.asarray(ref_idx, dtype=np.int32).reshape(n))
    states = np.empty((T, n, 9)) if want_states else None
    obs = np.empty((T, n, 5)) if want_obs else None
    rew = np.empty((T, n))
    done = np.empty((T, n), dtype=np.uint8)
    ep_len = np.empty(n, dtype=np.int32)
    xf = np.empty((n, 9))
    ret = np.empty(n)

    def p(a):
        return None if a is None else a.ctypes.data_as(C.c_void_p)

    total = lib().orc_rollout(n, T, dt, horizon, x0, actions, ref_bank, ref_idx, ref_bank.shape[1],
                              p(states), p(obs), p(rew), p(done), ep_len, xf, ret, nthreads)
    return dict(states=states, obs=obs, reward=rew, done=done, ep_len=ep_len, final_state=xf,
                total_return=ret, env_steps=int(total))


def trim_cost(V, H, u):
    """Cost.calculate (utils/find_trim_position.py:27-42)."""
    return lib().orc_trim_cost(float(V), float(H), _f64(u))


def trim(V, H, starts, nthreads=0):
    """Nelder-Mead trim search for every (cell, start): returns x [cells, S, 3], cost [cells, S], iters."""
    V, H = np.atleast_1d(_f64(V)), np.atleast_1d(_f64(H))
    starts = _f64(starts).reshape(-1, 3)
    n, S = V.size, starts.shape[0]
    x = np.empty((n, S, 3))
    cost = np.empty((n, S))
    iters = np.empty((n, S), dtype=np.int32)
    lib().orc_trim(n, V, H, S, starts, x, cost, iters.reshape(-1), nthreads)
    return x, cost, iters


def num_threads():
    return lib().orc_num_threads()
