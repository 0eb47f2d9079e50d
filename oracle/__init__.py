"""CPU oracle for the cardiax Fenton-Karma hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product path (``cardiax_b200``) never
imports this package and fails loudly when the CUDA library is missing.

PARITY UNPINNED (values): the reference (``/root/reference``) needs a 2021
JAX that is not installable here and its own tests hold no numerical golden
vectors for u/v/w.  The oracle is pinned on the only known-answer vector the
reference has (the stimulus schedule of ``tests/macro/stimulate_test.py:16-19``)
and on the ``tests/unittests/stimulus_test.py:23`` expectation; see DESIGN.md.
"""
from . import fk_oracle_ext  # noqa: F401  (odeint / resize / electrogram restatements)
from .fk_oracle import (  # noqa: F401
    Params, State, Protocol, Stimulus,
    init, gradient, stimulate, stimulus_active, stimulus_active_typed, step, step_euler, forward_euler, step_heun, forward_heun,
    tanh_xla_f32, PARAMSETS,
    rectangular, linear, triangular,
)
