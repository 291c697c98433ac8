"""Load the UNMODIFIED reference (`/root/reference/src`) behind the third-party stand-ins in oracle/shims.

TEST INFRASTRUCTURE ONLY — imported by oracle/gen_golden.py and by tests that pin the oracle ports against
the live reference; it only works in the authoring container (there is no /root/reference on the GPU box).
"""
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("TARL_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "simulation_core_model.py"))


class _RefModules:
    """Context manager that puts the reference's `src` package (and the shims) first on sys.path, evicting this
    repo's own drop-in `src` mirror from sys.modules for the duration, and restoring it afterwards."""

    def __enter__(self):
        self._saved_path = list(sys.path)
        self._saved_mods = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
        for k in self._saved_mods:
            del sys.modules[k]
        sys.path.insert(0, _SHIMS)
        sys.path.insert(0, REFERENCE_ROOT)
        return self

    def __exit__(self, *exc):
        self.ref_mods = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
        for k in self.ref_mods:
            del sys.modules[k]
        sys.modules.update(self._saved_mods)
        sys.path[:] = self._saved_path
        return False


_CACHE = {}


def load(*names):
    """Return reference modules by dotted name, e.g. load('src.simulation_core_model')."""
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    out = []
    missing = [n for n in names if n not in _CACHE]
    if missing:
        with _RefModules():
            # keep shim third-party packages importable by the already-loaded reference modules
            for n in missing:
                _CACHE[n] = importlib.import_module(n)
            for k, v in list(sys.modules.items()):
                if k == "src" or k.startswith("src."):
                    _CACHE.setdefault(k, v)
    for n in names:
        out.append(_CACHE[n])
    return out[0] if len(out) == 1 else out
