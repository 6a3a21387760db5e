"""Restatement of the one class the reference imports from ``ambiance==1.3.1``
(requirements.txt:2; call sites F16model/data/atmosphere.py:6-7,10-11).

TEST INFRASTRUCTURE ONLY. ``ambiance`` is a third-party dependency that is not
vendored under /root/reference and is not installable in this image (no
network, not in /opt/wheelhouse).  This module restates its published
algorithm -- the ICAO 1993 standard atmosphere (ICAO Doc 7488/3) -- so that the
UNMODIFIED reference can be imported by ``tests/golden/make_golden.py``.

PARITY UNPINNED at this boundary: the reference holds no test that pins
density / speed of sound.  Indirect pins: the two published trim points
(env/env.py:248-257 and control/ppo/pretty_plots.ipynb cell 13) only balance
if rho(h) and a(h) are right to ~1e-4 relative; both lie in the 0-11 km layer,
whose constants (p_b=101325, T_b=288.15, beta=-0.0065) are exact ICAO primary
constants.  Layers above 11 km use ambiance's tabulated, 6-digit p_b values,
restated from memory.
"""
import numpy as np


class CONST:
    g_0 = 9.80665
    R = 287.05287
    kappa = 1.4
    r = 6_356_766.0
    h_min = -5_004.0
    h_max = 81_020.0
    # (H_b [m], T_b [K], beta [K/m], p_b [Pa])
    LAYERS = np.array([
        [-5.0e3, 320.65, -6.5e-3, 1.77687e+5],
        [0.0e3, 288.15, -6.5e-3, 1.01325e+5],
        [11.0e3, 216.65, 0.0e-3, 2.26320e+4],
        [20.0e3, 216.65, 1.0e-3, 5.47487e+3],
        [32.0e3, 228.65, 2.8e-3, 8.68014e+2],
        [47.0e3, 270.65, 0.0e-3, 1.10906e+2],
        [51.0e3, 270.65, -2.8e-3, 6.69384e+1],
        [71.0e3, 214.65, -2.0e-3, 3.95639e+0],
    ])
    H_top = 80.0e3


class Atmosphere:
    def __init__(self, h, check_bounds=True):
        self.h = np.atleast_1d(np.asarray(h, dtype=float))
        if check_bounds and (np.any(self.h < CONST.h_min) or np.any(self.h > CONST.h_max)):
            raise ValueError(
                f"Value out of bounds. Lower limit: {CONST.h_min:.0f} m. "
                f"Upper limit: {CONST.h_max:.0f} m."
            )
        self.H = CONST.r * self.h / (CONST.r + self.h)
        idx = np.searchsorted(CONST.LAYERS[:, 0], self.H, side="right") - 1
        idx = np.clip(idx, 0, len(CONST.LAYERS) - 1)
        self._H_b, self._T_b, self._beta, self._p_b = CONST.LAYERS[idx].T

    @property
    def temperature(self):
        return self._T_b + self._beta * (self.H - self._H_b)

    @property
    def pressure(self):
        H, H_b, T_b, beta, p_b = self.H, self._H_b, self._T_b, self._beta, self._p_b
        T = self.temperature
        p_zero = p_b * np.exp((-CONST.g_0 / (CONST.R * T)) * (H - H_b))
        exponent = np.divide(1.0, beta, out=np.zeros_like(beta), where=beta != 0)
        p_nonzero = p_b * (1 + (beta / T_b) * (H - H_b)) ** (exponent * (-CONST.g_0 / CONST.R))
        return np.where(beta == 0, p_zero, p_nonzero)

    @property
    def density(self):
        return self.pressure / (CONST.R * self.temperature)

    @property
    def speed_of_sound(self):
        return np.sqrt(CONST.kappa * CONST.R * self.temperature)
