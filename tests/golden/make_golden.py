"""Generates tests/golden/*.npz with the CPU oracle (NumPy restatement of the reference).

The reference itself cannot be imported here (JAX absent), so these are ORACLE
vectors, not reference vectors: they pin the CUDA path and the C port to the literal NumPy
restatement.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle as O  # noqa: E402
from tests import common  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def scar_s1s2():
    """A small version of BASELINE config 2: heterogeneous D, S1 (NORTH) - S2 (WEST, rotated) cross-field."""
    shape = (64, 96)
    st, D = common.smooth_case(shape, seed=1)
    s1 = O.triangular(shape, 0, 10.0, 0.2, 20.0, O.Protocol(0, 2, 1e9))
    s2 = O.triangular(shape, 0, 100.0, 0.5, 20.0, O.Protocol(40, 2, 1e9))
    params = O.PARAMSETS["3"]
    cps = np.arange(0, 121, 30)
    out = {"D": D, "params": np.array(params, np.float64), "checkpoints": cps, "dt": 0.01, "dx": 0.01,
           "v0": st.v, "w0": st.w, "u0": st.u}
    for i, s in enumerate((s1, s2)):
        out["field%d" % i] = s.field
        out["proto%d" % i] = np.array(s.protocol, np.float64)
    s = st
    for i in range(len(cps) - 1):
        s = O.forward_euler(s, cps[i], cps[i + 1], params, D, [s1, s2], 0.01, 0.01)
        out["v%d" % (i + 1)], out["w%d" % (i + 1)], out["u%d" % (i + 1)] = s
    np.savez_compressed(os.path.join(HERE, "fk_64x96_scar_s1s2.npz"), **out)


def wave_128():
    """BASELINE config 1 (README benchmark shape): summary values of the 128^2 plane wave at 1e3 steps."""
    shape = (128, 128)
    st = O.init(shape)
    D = np.full(shape, 1e-3, np.float32)
    stim = [O.linear(shape, 0, 0.2, 20.0, O.Protocol(0, 2, 1e9))]
    from oracle import c_oracle
    s = c_oracle.forward_euler(st, 0, 1000, O.PARAMSETS["3"], D, stim, 0.01, 0.01)
    np.savez_compressed(os.path.join(HERE, "fk_128_wave_1000.npz"), v=s.v[::4, ::4], w=s.w[::4, ::4], u=s.u[::4, ::4],
                        u_col=s.u[:, 64], v_col=s.v[:, 64], w_col=s.w[:, 64])


if __name__ == "__main__":
    scar_s1s2()
    wave_128()
    print(os.listdir(HERE))
