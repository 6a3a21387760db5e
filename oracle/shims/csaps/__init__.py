"""Stub for ``csaps`` (imported at F16model/utils/interpolator.py:3, used only by
the ``interp_method="csaps"`` branch, which the hot path never selects).
TEST INFRASTRUCTURE ONLY."""


def csaps(*args, **kwargs):  # pragma: no cover
    raise NotImplementedError("csaps is not available; the hot path uses method='linear'")
