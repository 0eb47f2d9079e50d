"""Generates tests/golden/fk_dopri_24x28.npz with the CPU oracle's restatement of jax.experimental.ode.odeint
(oracle/fk_oracle_ext.py).  ORACLE vectors (jax cannot be imported here: PARITY UNPINNED).
Run from the repo root:  python tests/golden/make_dopri.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle as O  # noqa: E402
from oracle import fk_oracle_ext as X  # noqa: E402
from tests import common  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SHAPE = (24, 28)
TS = np.array([0, 0.5, 1.0, 2.5], np.float32)


def case():
    st, D = common.smooth_case(SHAPE, 1)
    stim = [O.linear(SHAPE, 0, 0.3, 20.0, O.Protocol(0, 2, 50))]
    return st, D, stim


if __name__ == "__main__":
    st, D, stim = case()
    out = {"ts": TS}
    for name, tol in (("tight", 1.4e-8), ("loose", 1e-5)):
        stats = {}
        ref = X.odeint_dopri5(st, TS, O.PARAMSETS["3"], D, stim, 0.01, rtol=tol, atol=tol, stats=stats)
        out["v_" + name], out["w_" + name], out["u_" + name] = ref
        out["stats_" + name] = np.array([stats["attempts"], stats["accepted"], stats["rhs_evals"]], np.int64)
        out["tol_" + name] = tol
    np.savez_compressed(os.path.join(HERE, "fk_dopri_24x28.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
