"""Drop-in ``cardiax`` namespace: the reference's module names bound to the B200 implementation.

``import cardiax; cardiax.solve.forward(...)`` and ``from cardiax import params, stimulus`` work as
with the reference (cardiax/__init__.py:1); plot/io/metrics are imported lazily because their
third-party dependencies are optional.
"""
import sys as _sys

from cardiax_b200 import convert, options, params, solve, stimulus  # noqa: F401

for _n in ("convert", "params", "solve", "stimulus"):
    _sys.modules[__name__ + "." + _n] = globals()[_n]


def __getattr__(name):
    if name in ("io", "plot", "metrics"):
        import importlib
        mod = importlib.import_module("cardiax_b200." + name)
        _sys.modules[__name__ + "." + name] = mod
        return mod
    raise AttributeError(name)
