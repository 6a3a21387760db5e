def __getattr__(name):  # pragma: no cover
    raise NotImplementedError("matplotlib is stubbed out in the oracle harness")
