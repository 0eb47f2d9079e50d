"""BASELINE config 2 at its REAL protocol, frozen: 512 x 512 scar map, S1 at step 0, S2 (rotated by 90 degrees) at step
40 000 (experiments/generate_fd_data_256.py:22-40, 11-12: segments of 500 steps), PARAMSET_3, dt = dx = 0.01 -- the run up
to step 40 500, i.e. ACROSS the second stimulus, computed by the C oracle (bit-identical to the NumPy restatement and,
through tests/test_reference_pin.py, to the reference's own source).

Stored: SHA-256 of v, w, u at steps 39 500, 40 000 (just before S2), 40 002 (both S2 steps applied) and 40 500 for the
exact-numerics comparison; u, v, w subsampled 4 x 4 at 40 002 and 40 500 in float32 AND float64 arithmetic for the
fast-numerics envelope; hashes of the regenerated inputs.  About 3 minutes on 8 cores:

    python tests/golden/make_config2.py
"""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle as O  # noqa: E402
from oracle import c_oracle as C  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SHAPE = (512, 512)
MARKS = (39500, 40000, 40002, 40500)


def inputs(seed=0):
    """The bench's `make_fk512` inputs (bench.py), regenerated from the seed."""
    from cardiax_b200 import generate
    D = np.ascontiguousarray(generate.random_diffusivity(np.random.default_rng(seed), SHAPE), dtype=np.float32)
    ang = float(np.random.default_rng(seed).uniform(0, 180))
    s1 = O.triangular(SHAPE, 0, ang, 0.2, 20.0, O.Protocol(0, 2, 1e9))
    s2 = O.triangular(SHAPE, 0, ang + 90, 0.5, 20.0, O.Protocol(40000, 2, 1e9))
    return D, [s1, s2]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    D, stim = inputs()
    out = {"marks": np.array(MARKS), "D_sha": sha(D), "s1_sha": sha(stim[0].field), "s2_sha": sha(stim[1].field)}
    C.set_threads(os.cpu_count() or 1)
    for dtype, tag in ((np.float32, "f32"), (np.float64, "f64")):
        s, t = O.init(SHAPE, dtype), 0
        for m in MARKS:
            s = C.forward_euler(s, t, m, O.PARAMSETS["3"], D, stim, 0.01, 0.01, dtype=dtype)
            t = m
            if tag == "f32":
                out["sha_%d" % m] = np.array([sha(x) for x in s])
            if m >= 40002:
                for name, x in zip("vwu", s):
                    out["%s_%s_%d" % (name, tag, m)] = np.ascontiguousarray(x[::4, ::4])
            print(tag, m, float(s.u.max()), flush=True)
    np.savez_compressed(os.path.join(HERE, "fk_512_config2_s2.npz"), **out)


if __name__ == "__main__":
    main()
