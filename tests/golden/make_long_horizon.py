"""Generates tests/golden/fk_192_breakup.npz: a long, chaotic spiral-breakup run of the CPU oracle (C port, bit-identical
to the NumPy restatement) reduced to the observables the long-horizon parity tests compare.

Case: 192 x 192, PARAMSET_5 (the reference's "break-up" regime, experiments/generate_fd_examples.py:47-51), the
reference's dt = dx = 0.01, D = 1e-4 (the lower end of the reference's scar maps, deepx/generate.py:80: it shrinks the
wave so that spirals fit 192 cells), no stimulus, initial state = a smoothed broken plane wave (tests/analysis.py) that
curls into a spiral and breaks up.  36 snapshots 1000 steps apart.

Stored: the activation-time maps and spiral-tip trajectories of the fp32 oracle, the SHA-256 of its state at steps 4000
and 36000 (exact-numerics kernels must reproduce them bit for bit), and the ENVELOPE: the same observables' deviation
between the fp32 oracle and its fp64 twin, which is the yardstick for the fast-numerics kernels.

    python tests/golden/make_long_horizon.py          (about two minutes on 8 cores)
"""
import hashlib
import os
import sys

import numpy as np
from scipy import ndimage

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle as O  # noqa: E402
from oracle import c_oracle as C  # noqa: E402
from tests import analysis as An  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SHAPE, PSET, DVAL, DT, DX, NSEG, SEG = (192, 192), "5", 1e-4, 0.01, 0.01, 36, 1000


def initial_state():
    return [ndimage.gaussian_filter(x, 2.0).astype(np.float32) for x in An.broken_wave(SHAPE, width=12)]


def state_hash(st):
    h = hashlib.sha256()
    for x in st:
        h.update(np.ascontiguousarray(x, dtype=np.float32).tobytes())
    return h.hexdigest()


def run(dtype):
    D = np.full(SHAPE, DVAL, dtype)
    st = O.State(*[x.astype(dtype) for x in initial_state()])
    frames, hashes = [np.array(st.u, np.float32)], {}
    for k in range(NSEG):
        st = C.forward_euler(st, k * SEG, (k + 1) * SEG, O.PARAMSETS[PSET], D, [], DT, DX, dtype=dtype)
        frames.append(np.array(st.u, np.float32))
        if dtype == np.float32 and (k + 1) in (4, NSEG):
            hashes[k + 1] = state_hash(st)
    return np.stack(frames), hashes


def main():
    C.set_threads(os.cpu_count() or 1)
    f32, hashes = run(np.float32)
    f64, _ = run(np.float64)
    act32, act64 = An.activation_times(f32, beats=2), An.activation_times(f64, beats=2)
    tips32, tips64 = An.tip_trajectory(f32), An.tip_trajectory(f64)
    env_tip = An.tip_distance(tips32, tips64)
    env_maxabs = np.array([np.abs(f32[k] - f64[k]).max() for k in range(NSEG + 1)])
    env_act = np.abs(act32 - act64)
    # horizon: the snapshot pairs before the fp32 and fp64 oracles' tips first part by more than 2 cells
    bad = np.nonzero(~(np.nan_to_num(env_tip, nan=0.0) <= 2.0))[0]
    horizon = int(bad[0]) if len(bad) else NSEG
    out = {
        "shape": np.array(SHAPE), "pset": PSET, "D": DVAL, "dt": DT, "dx": DX, "nseg": NSEG, "seg": SEG,
        "act32": act32.astype(np.float32), "env_act": env_act.astype(np.float32),
        "tips32": np.array([np.pad(t, ((0, 16 - len(t)), (0, 0)), constant_values=np.nan) for t in tips32], np.float32),
        "ntips32": np.array([len(t) for t in tips32]), "env_tip": env_tip, "env_maxabs": env_maxabs,
        "horizon": horizon, "hash_step4000": hashes[4], "hash_final": hashes[NSEG],
        "u_final_crop": f32[-1][::8, ::8],
    }
    np.savez_compressed(os.path.join(HERE, "fk_192_breakup.npz"), **out)
    print("horizon (snapshot pairs):", horizon, " tips per pair:", out["ntips32"].tolist())
    print("env maxabs:", np.round(env_maxabs, 4).tolist())
    print("env tip distance:", np.round(env_tip, 2).tolist())


if __name__ == "__main__":
    main()
