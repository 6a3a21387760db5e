"""Model layer -- batched mirror of ``F16model/model`` (``States``, ``Control``,
``F16model.step``, ``ODE_3DoF.solve``), backed by the CUDA library.

``States`` / ``Control`` keep the reference's field names and arithmetic
(model/_interface.py:6-90) but each field may be a scalar or a tensor of N values."""

import numpy as np

from . import _capi

STATE_FIELDS = ("Ox", "Oy", "wz", "theta", "V", "alpha", "stab", "dstab", "Pa")


class States:
    def __init__(self, Ox, Oy, wz, theta, V, alpha, stab, dstab, Pa):
        self.Ox, self.Oy, self.wz, self.theta, self.V = Ox, Oy, wz, theta, V
        self.alpha, self.stab, self.dstab, self.Pa = alpha, stab, dstab, Pa

    def to_array(self):
        return np.array([getattr(self, f) for f in STATE_FIELDS])

    @classmethod
    def from_array(cls, arr):
        return cls(*[arr[i] for i in range(9)])

    def __add__(self, other):
        if isinstance(other, States):
            return States(*[getattr(other, f) + getattr(self, f) for f in STATE_FIELDS])
        return NotImplemented

    def __rmul__(self, other):
        if np.isscalar(other):
            return States(*[getattr(self, f) * other for f in STATE_FIELDS])
        if isinstance(other, States):
            return States(*[getattr(self, f) * getattr(other, f) for f in STATE_FIELDS])
        return NotImplemented

    __mul__ = __rmul__

    def __repr__(self):
        return (f"Ox = {self.Ox} m; Oy = {self.Oy} m; wz = {np.degrees(self.wz)} deg/s; V = {self.V} m/s; "
                f"theta = {np.degrees(self.theta)} deg; stab_pos = {np.degrees(self.stab)} deg; "
                f"dstab = {np.degrees(self.dstab)} deg/s; thrust = {self.Pa} H %")


class Control:
    def __init__(self, stab, throttle):
        self.stab = stab          # rad
        self.throttle = throttle  # 0..1

    def to_array(self):
        return np.array([self.stab, self.throttle])

    def __rmul__(self, other):
        if np.isscalar(other):
            return Control(self.stab * other, self.throttle * other)
        return NotImplemented

    __mul__ = __rmul__

    def __repr__(self):
        return f"stab = {np.degrees(self.stab)} deg; throttle = {self.throttle};"


def _as_soa(x, rows, name):
    """Accept [rows, n] (SoA, native) or [n, rows]; return a contiguous [rows, n] CUDA tensor."""
    import torch

    if not isinstance(x, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor on a CUDA device")
    if x.dim() == 1:
        x = x.reshape(rows, 1)
    if x.shape[0] != rows:
        if x.shape[-1] != rows:
            raise ValueError(f"{name} must have shape [{rows}, n] or [n, {rows}]")
        x = x.t()
    return _capi.require_cuda(x.contiguous(), name)


def rhs(x, u):
    """``ODE_3DoF.solve`` for a batch: x [9, n], u [2, n] (stab command [rad], throttle) ->
    derivatives [9, n] including V_dot (model/ODE_3DoF.py:8-60)."""
    import torch

    x = _as_soa(x, 9, "x")
    u = _as_soa(u, 2, "u").to(x.dtype)
    if u.shape[1] != x.shape[1]:
        raise ValueError("x and u batch sizes differ")
    dx = torch.empty_like(x)
    L = _capi.load_library()
    with torch.cuda.device(x.device):
        _capi.check(L.f16_rhs(_capi.dtype_code(x.dtype), x.shape[1], x.data_ptr(), u.data_ptr(), dx.data_ptr(),
                              _capi.stream_ptr(x.device)))
    return dx


def model_step(x, u, dt, steps=1):
    """``steps`` x ``F16model.step`` with the control held: explicit Euler, wz clipped to
    +-60 deg/s, V frozen (model/_interface.py:116-122).  x [9, n] is advanced IN PLACE."""
    import torch

    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.is_contiguous() and x.dim() == 2 and x.shape[0] == 9):
        raise ValueError("x must be a contiguous CUDA tensor of shape [9, n]")
    u = _as_soa(u, 2, "u").to(x.dtype)
    L = _capi.load_library()
    with torch.cuda.device(x.device):
        _capi.check(L.f16_model_step(_capi.dtype_code(x.dtype), x.shape[1], x.data_ptr(), u.data_ptr(), float(dt),
                                     int(steps), _capi.stream_ptr(x.device)))
    return x


class F16model:
    """Bare integrator interface (model/_interface.py:108-126) over a batch of n aircraft.
    ``x0`` is a ``States`` of scalars (n = 1) or a [9, n] CUDA tensor."""

    def __init__(self, x0, dt=0.001, device="cuda", dtype=None):
        import torch

        if isinstance(x0, States):
            dtype = dtype or torch.float64
            x = torch.tensor(np.asarray(x0.to_array(), dtype=np.float64).reshape(9, -1), dtype=dtype, device=device)
        else:
            x = _as_soa(x0, 9, "x0").clone()
        self.init_state = x.clone()
        self.state_prev = x
        self.dt = dt

    def step(self, u):
        import torch

        if isinstance(u, Control):
            u = torch.tensor(np.asarray(u.to_array(), dtype=np.float64).reshape(2, -1), dtype=self.state_prev.dtype,
                             device=self.state_prev.device).expand(2, self.state_prev.shape[1]).contiguous()
        model_step(self.state_prev, u, self.dt, 1)
        return self.state_prev

    def reset(self):
        self.state_prev = self.init_state.clone()
        return self.state_prev


# data-level probes -----------------------------------------------------------
def _vec_call(fn_name, dtype_ref, inputs, n_out_rows):
    import torch

    ins = [_capi.require_cuda(t.contiguous().to(dtype_ref.dtype), "input") for t in inputs]
    n = ins[0].numel()
    out = torch.empty((n_out_rows, n), dtype=dtype_ref.dtype, device=dtype_ref.device)
    L = _capi.load_library()
    with torch.cuda.device(dtype_ref.device):
        args = [_capi.dtype_code(dtype_ref.dtype), n] + [t.data_ptr() for t in ins]
        if fn_name == "f16_atmosphere":
            args += [out[0].data_ptr(), out[1].data_ptr()]
        else:
            args += [out.data_ptr()]
        args.append(_capi.stream_ptr(dtype_ref.device))
        _capi.check(getattr(L, fn_name)(*args))
    return out


def aero_coeffs(alpha, fi, wz, V):
    """(Cx, Cy, Mz) [3, n] -- coeff.get_Cx/get_Cy/get_Mz at beta=0, lef=0, sb=0."""
    return _vec_call("f16_coeffs", alpha, [alpha, fi, wz, V], 3)


def thrust(H, M, Pa):
    """thrust.get_thrust(H, M, Pa) [n] (data/thrust.py:6-14)."""
    return _vec_call("f16_thrust", H, [H, M, Pa], 1)[0]


def atmosphere(h):
    """(density, speed of sound) [2, n] (data/atmosphere.py:5-11)."""
    return _vec_call("f16_atmosphere", h, [h], 2)
