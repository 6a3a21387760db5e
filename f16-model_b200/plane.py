"""Airframe constants and bounds -- the host-side mirror of F16model/data/plane.py:3-34
(the device copy lives in csrc/f16_device.cuh, namespace K)."""
import numpy as np

m = 9295.44  # kg
l = 9.144  # m  # noqa: E741
S = 27.87  # m^2
b_a = 3.45  # m
Jz = 75673.6
rcgx = -0.05 * b_a

Tstab = 0.03
Xistab = 0.707
maxabsstab = np.radians(25)
maxabsdstab = np.radians(60)
minthrottle = 0
maxthrottle = 1

lef = 0
sb = 0

# dict order matters: it is the order of the first three observation entries (env/env.py:93-97)
state_bound = {
    "Oy": [0, 35000],
    "wz": [np.radians(-150), np.radians(150)],
    "V": [0, 500],
}

random_state_bound = {
    "Oy": [2000, 4000],
    "wz": [np.radians(-5), np.radians(5)],
    "V": [90, 250],
    "alpha": [np.radians(-5), np.radians(5)],
    "maxabsstab": [np.radians(-5), np.radians(5)],
}
