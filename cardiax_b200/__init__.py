"""cardiax_b200 -- B200-native Fenton-Karma tissue stepper behind the cardiax API.

``from cardiax_b200 import solve, params, stimulus, convert`` mirrors the reference's
``cardiax.solve`` / ``cardiax.params`` / ``cardiax.stimulus`` / ``cardiax.convert``; the top-level
``cardiax`` package in this repository re-exports the same modules so reference scripts run
unchanged.  The compute path is libfk.so (hand-written sm_100a CUDA behind a C ABI, include/fk.h).
"""
from . import convert, options, params, stimulus  # noqa: F401
from . import solve  # noqa: F401

__all__ = ["convert", "options", "params", "solve", "stimulus"]
