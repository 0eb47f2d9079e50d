"""Freezes outputs of the UNMODIFIED reference (``/root/reference/cardiax``, run on ``ref_shim``'s NumPy stand-in for
jax) as ``tests/golden/ref_*.npz``, so that machines without the reference tree (the GPU box) still compare against
vectors the reference's own source produced.  Run from the repo root, in the build container:

    python tests/golden/make_reference_golden.py

``tanh`` = "xla" vectors use XLA's published fp32 tanh (what jaxlib 0.1.64 would emit); "numpy" vectors use np.tanh.
The CUDA path's exact numerics and the oracle's default both implement the "xla" function.
"""
import io
import os
import sys
from contextlib import redirect_stdout

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import common  # noqa: E402
from tests.golden import ref_shim as S  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
PARAMSET_NAMES = ["1A", "1B", "1C", "1D", "1E", "2", "3", "4A", "4B", "4C", "5", "6", "7", "8", "9", "10"]


def _stimuli(ref, shape):
    RS = ref.stimulus
    return [RS.linear(shape, RS.Direction.NORTH, 0.25, 20.0, RS.Protocol(0, 2, 1e9)),
            RS.triangular(shape, RS.Direction.WEST, 100.0, 0.5, 20.0, RS.Protocol(40, 2, 1e9)),
            RS.rectangular(shape, (shape[0] // 2, shape[1] // 3), (10, 8), -5.0, RS.Protocol(3, 3, 17))]


def integrators(ref):
    """solve.forward (Euler and Heun, float32 counter) + _forward_euler with an int32 counter on a small scar tissue."""
    shape = (48, 40)
    st, D = common.smooth_case(shape, seed=11)
    stim = _stimuli(ref, shape)
    cps = np.arange(0, 121, 30)
    out = {"D": D, "v0": st.v, "w0": st.w, "u0": st.u, "checkpoints": cps, "dt": 0.01, "dx": 0.01,
           "params": np.array(ref.params.PARAMSET_3, np.float64), "sha256": np.array(sorted(S.file_hashes().items()))}
    for i, s in enumerate(stim):
        out["field%d" % i] = S.to_numpy(s.field)
        out["proto%d" % i] = np.array([float(np.asarray(x).reshape(-1)[0]) for x in s.protocol], np.float64)
    for tanh in ("xla", "numpy"):
        with ref.tanh(tanh), redirect_stdout(io.StringIO()):
            for name, integ in (("euler", ref.solve.TimeIntegrator.EULER), ("heun", ref.solve.TimeIntegrator.HEUN)):
                states = ref.solve.forward(ref.solve.State(*st), cps, ref.params.PARAMSET_3, D, stim, 0.01, 0.01, integ)
                for i, s in enumerate(states):
                    for f, a in zip("vwu", S.to_numpy(s)):
                        out["%s_%s_%s%d" % (name, tanh, f, i + 1)] = a
            s = ref.solve._forward_euler(ref.solve.State(*st), 0, 60, ref.params.PARAMSET_3, D, stim, 0.01, 0.01)
            for f, a in zip("vwu", S.to_numpy(s)):
                out["euler_int_%s_%s" % (tanh, f)] = a
    np.savez_compressed(os.path.join(HERE, "ref_fk_48x40.npz"), **out)


def steps(ref):
    """solve.step for all 16 parameter sets (white-noise state around the gate thresholds, heterogeneous D, an active
    stimulus), solve.gradient on a 4-D array, solve.stimulate over a schedule."""
    shape = (21, 26)
    st, D, stim = common.random_case(shape, seed=3, n_stim=3)
    rs = [ref.stimulus.Stimulus(ref.stimulus.Protocol(*s.protocol), s.field) for s in stim]
    out = {"D": D, "v0": st.v, "w0": st.w, "u0": st.u, "dx": 0.01, "t": 4.0}
    for i, s in enumerate(stim):
        out["field%d" % i] = s.field
        out["proto%d" % i] = np.array(s.protocol, np.float64)
    for tanh in ("xla", "numpy"):
        with ref.tanh(tanh):
            for key in PARAMSET_NAMES:
                d = S.to_numpy(ref.solve.step(ref.solve.State(*st), 4.0, getattr(ref.params, "PARAMSET_" + key), D, rs, 0.01))
                for f, a in zip("vwu", d):
                    out["d%s_%s_%s" % (f, tanh, key)] = a
    a = np.random.default_rng(0).standard_normal((5, 6, 7, 9)).astype(np.float32)
    out["grad_in"] = a
    for axis in (0, 1, 2, 3):
        out["grad_axis%d" % axis] = S.to_numpy(ref.solve.gradient(a, axis))
    X = np.random.default_rng(1).standard_normal(shape).astype(np.float32)
    out["stimulate_in"] = X
    out["stimulate_out"] = np.stack([S.to_numpy(ref.solve.stimulate(float(t), X, rs)) for t in range(24)])
    np.savez_compressed(os.path.join(HERE, "ref_fk_steps.npz"), **out)


if __name__ == "__main__":
    ref = S.load_reference()
    assert S.file_hashes() == S.REFERENCE_SHA256, "reference files changed: re-validate the stand-in first"
    integrators(ref)
    steps(ref)
    print(sorted(f for f in os.listdir(HERE) if f.startswith("ref_")))
