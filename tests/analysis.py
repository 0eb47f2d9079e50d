"""Wave-level observables used by the long-horizon parity tests (test infrastructure, NumPy only).

The reference defines none of these (cardiax/metrics.py:25-26 `spiral_centres` is a stub), so the definitions are the
usual ones of the cardiac-simulation literature:

* activation time  -- the (linearly interpolated) time at which u crosses 0.5 upwards, per cell and per beat;
* spiral tip       -- a point where the u = 0.5 isolines of two consecutive snapshots intersect, i.e. u = 0.5 and
                      du/dt = 0 (Fenton & Karma 1998, section III).
"""
import numpy as np
from scipy import ndimage


def broken_wave(shape, width=10):
    """Initial state that curls into a single spiral: a plane wave front covering the upper half of the tissue only,
    with refractory tissue (closed v, w gates) in its wake.  Returns (v, w, u) float32."""
    H, W = shape
    u = np.zeros(shape, np.float32)
    v = np.ones(shape, np.float32)
    w = np.ones(shape, np.float32)
    h, c = H // 2, W // 2
    u[:h, c - width:c] = 1.0
    v[:h, :c - width] = 0.0
    w[:h, :c - width] = 0.0
    return v, w, u


def activation_times(frames, threshold=0.5, beats=3):
    """frames: (F, H, W) snapshots of u.  Returns (beats, H, W): time, in snapshot units, of the k-th upward crossing of
    `threshold` in every cell (NaN where there is none)."""
    frames = np.asarray(frames, dtype=np.float64)
    F, H, W = frames.shape
    out = np.full((beats, H, W), np.nan)
    count = np.zeros((H, W), np.int64)
    for f in range(F - 1):
        a, b = frames[f], frames[f + 1]
        cross = (a < threshold) & (b >= threshold)
        if not cross.any():
            continue
        t = f + (threshold - a) / np.where(cross, b - a, 1.0)
        for k in range(beats):
            sel = cross & (count == k)
            out[k][sel] = t[sel]
        count += cross
    return out


def spiral_tips(u_a, u_b, threshold=0.5):
    """Tips between two consecutive snapshots: centroids (row, col) of the connected groups of 2 x 2 cell blocks in
    which BOTH u_a - threshold and u_b - threshold change sign."""
    def straddles(x):
        s = x >= threshold
        q = s[:-1, :-1].astype(np.int8) + s[1:, :-1] + s[:-1, 1:] + s[1:, 1:]
        return (q > 0) & (q < 4)
    cand = straddles(np.asarray(u_a)) & straddles(np.asarray(u_b))
    if not cand.any():
        return np.zeros((0, 2))
    lab, n = ndimage.label(ndimage.binary_dilation(cand, iterations=1))
    lab = lab * cand
    tips = ndimage.center_of_mass(cand, lab, index=np.arange(1, n + 1))
    return np.array([(r + 0.5, c + 0.5) for r, c in tips if np.isfinite(r)])


def tip_trajectory(frames, threshold=0.5, margin=3):
    """List (one entry per consecutive snapshot pair) of tip arrays, tips closer than `margin` cells to the tissue
    edge dropped (a wave front ending on the boundary is not a spiral tip)."""
    out = []
    H, W = frames[0].shape
    for f in range(len(frames) - 1):
        t = spiral_tips(frames[f], frames[f + 1], threshold)
        if len(t):
            keep = (t[:, 0] > margin) & (t[:, 0] < H - margin) & (t[:, 1] > margin) & (t[:, 1] < W - margin)
            t = t[keep]
        out.append(t)
    return out


def tip_distance(traj_a, traj_b):
    """Per snapshot pair: max over the tips of A of the distance to the nearest tip of B (NaN when either has none;
    inf when only one of them has tips)."""
    d = []
    for a, b in zip(traj_a, traj_b):
        if len(a) == 0 and len(b) == 0:
            d.append(np.nan)
        elif len(a) == 0 or len(b) == 0:
            d.append(np.inf)
        else:
            dist = np.sqrt(((a[:, None, :] - b[None, :, :]) ** 2).sum(-1))
            d.append(max(dist.min(1).max(), dist.min(0).max()))
    return np.array(d)
