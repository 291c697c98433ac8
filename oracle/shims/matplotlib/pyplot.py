def __getattr__(name):
    raise RuntimeError("matplotlib is not installed; plotting is outside the oracle's scope")
