"""CPU oracle for the cardiax Fenton-Karma hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product path (``cardiax_b200``) never
imports this package and fails loudly when the CUDA library is missing.

PINNED TO THE REFERENCE'S OWN SOURCE: ``tests/test_reference_pin.py`` imports the
unmodified ``/root/reference/cardiax/{solve,stimulus,params,convert}.py`` by path on
a NumPy stand-in for jax (``tests/golden/ref_shim.py``: jax's 32-bit promotion
lattice, weak Python scalars, ``fori_loop`` counter dtype, ``jnp.mod``) and asserts
bit-equality with this package for ``step``, ``step_euler``, ``_forward_euler``,
``step_heun``, ``_forward_heun``, N-D ``gradient``, ``stimulate``, the three mask
builders, all 16 parameter sets, the float32 and the int32 loop counter; vectors
frozen from those runs (``tests/golden/ref_fk_*.npz``) travel to the GPU box.
Still unpinnable here: what XLA does below the op level (FMA contraction, its
``tanh``: restated from the published rational, ``tanh="xla"``), and jax's
``odeint`` (``fk_oracle_ext``: PARITY UNPINNED).
"""
from . import fk_oracle_ext  # noqa: F401  (odeint / resize / electrogram restatements)
from .fk_oracle import (  # noqa: F401
    Params, State, Protocol, Stimulus,
    init, gradient, stimulate, stimulus_active, stimulus_active_typed, step, step_euler, forward_euler, step_heun, forward_heun,
    tanh_xla_f32, PARAMSETS,
    rectangular, linear, triangular,
)
