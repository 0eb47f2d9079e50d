"""Shared seeded test inputs (NumPy, host)."""
import numpy as np

import oracle as O


def random_case(shape, seed=0, n_stim=2, hetero=True):
    """Random but physiological state, heterogeneous diffusivity in [1e-4, 1e-3], two stimuli."""
    rng = np.random.default_rng(seed)
    st = O.State(rng.random(shape, dtype=np.float32), rng.random(shape, dtype=np.float32),
                 (rng.random(shape, dtype=np.float32) * 1.2 - 0.1).astype(np.float32))
    if hetero:
        D = (rng.random(shape, dtype=np.float32) * 9e-4 + 1e-4).astype(np.float32)
    else:
        D = np.full(shape, 1e-3, np.float32)
    stim = [O.linear(shape, 0, 0.3, 20.0, O.Protocol(0, 2, 50)),
            O.rectangular(shape, (shape[0] // 2, shape[1] // 2), (10, 10), -5.0, O.Protocol(3, 2, 7)),
            O.triangular(shape, 3, 30.0, 0.4, 0.7, O.Protocol(1, 3, 11))][:n_stim]
    return st, D, stim


def smooth_case(shape, seed=0):
    """A smooth excitable field (blobs of excitation) -- closer to a real run than white noise."""
    rng = np.random.default_rng(seed)
    H, W = shape
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    u = np.zeros(shape, np.float32)
    for _ in range(6):
        cy, cx, r = rng.uniform(0, H), rng.uniform(0, W), rng.uniform(4, max(5, min(H, W) / 6))
        u += np.exp(-(((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * r * r))).astype(np.float32)
    u = np.clip(u, 0, 1).astype(np.float32)
    v = (1 - 0.8 * u).astype(np.float32)
    w = (1 - 0.3 * u).astype(np.float32)
    D = (1e-4 + 9e-4 * (0.5 + 0.5 * np.sin(xx / 7.0) * np.cos(yy / 9.0))).astype(np.float32)
    return O.State(v, w, u), D


def to_oracle_stimuli(stimuli):
    return [O.Stimulus(O.Protocol(*s.protocol), np.asarray(s.field.cpu() if hasattr(s.field, "cpu") else s.field))
            for s in stimuli]
