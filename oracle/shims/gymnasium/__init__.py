"""Minimal stand-in for ``gymnasium==0.29.1`` (F16model/env/env.py:3): only the
names the reference env touches -- ``Env.reset(seed=)`` and ``spaces.Box``.
TEST INFRASTRUCTURE ONLY; contributes no arithmetic."""
import numpy as np
from . import spaces  # noqa: F401


class Env:
    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self._np_random = np.random.default_rng(seed)
