"""Drop-in ``deepx`` namespace -- ONLY the data-generation driver (deepx/generate.py), which is the caller of the
accelerated solver path.  The neural surrogate (resnet, optimise, dataset) is out of scope."""
import sys as _sys

from cardiax_b200 import generate  # noqa: F401

_sys.modules[__name__ + ".generate"] = generate
