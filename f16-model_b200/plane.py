"""Airframe constants and bounds of the longitudinal F-16 model, host-side copy.

The authoritative copy used by the kernels lives in ``csrc/f16_device.cuh`` (namespace ``K``);
this module exposes the same quantities under the names the reference's ``F16model.data.plane``
uses (F16model/data/plane.py:3-34), because the env helpers (``normalize``, ``rescale_action``,
``get_random_state``) and user code read them by those names."""
import math as _math

_DEG = _math.pi / 180.0


def _deg_range(lo, hi):
    return [lo * _DEG, hi * _DEG]


# mass / geometry (SI units)
m = 9295.44         # mass [kg]
l = 9.144           # span [m] (lateral model only)  # noqa: E741 (the reference's name)
S = 27.87           # wing area [m^2]
b_a = 3.45          # mean aerodynamic chord [m]
Jz = 75673.6        # pitch inertia [kg m^2]
rcgx = -0.05 * b_a  # centre-of-gravity offset [m]

# stabilator actuator (second-order lag with rate limit) and throttle limits
Tstab = 0.03
Xistab = 0.707
maxabsstab = 25 * _DEG
maxabsdstab = 60 * _DEG
minthrottle = 0
maxthrottle = 1

# configuration held fixed by the longitudinal model: leading-edge flap and speed brake retracted
lef = 0
sb = 0

# Observation normalisation ranges.  The ORDER of the keys is the order of the first three observation
# entries (env/env.py:93-97): altitude, pitch rate, airspeed.
state_bound = dict(Oy=[0, 35000], wz=_deg_range(-150, 150), V=[0, 500])

# Box the random initial state is drawn from (get_random_state, env/env.py:221-243)
random_state_bound = dict(Oy=[2000, 4000], wz=_deg_range(-5, 5), V=[90, 250], alpha=_deg_range(-5, 5),
                          maxabsstab=_deg_range(-5, 5))
