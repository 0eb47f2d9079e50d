"""cardiax/metrics.py on the GPU.  Like the reference, only ``electrogram`` does anything: ``adp``, ``restitution`` and
``spiral_centres`` are ``pass`` stubs there (metrics.py:5-10, 25-26) and return None here too."""
import ctypes

import torch

from . import _lib, solve


def adp(x, perc):
    """cardiax/metrics.py:5-6 -- a stub in the reference."""
    return None


def restitution(x, perc):
    """cardiax/metrics.py:9-10 -- a stub in the reference."""
    return None


def electrogram(x, point):
    """cardiax/metrics.py:13-22 -- ``sum(x * dist, axis=(-1, -2))`` with ``dist[i, j] = sqrt((j - point[0])**2 +
    (i - point[1])**2)``: one value per (H, W) frame of ``x`` (leading axes kept).  The reference builds ``dist`` from
    ``ogrid[:W, :H]``, which only broadcasts against square frames; the same restriction is enforced here.  (The
    reference also prints ``dist``; that debugging print is not reproduced.)"""
    x = solve._as_f32(x)
    if x.dim() < 2:
        raise ValueError("electrogram expects frames of shape (..., H, W)")
    H, W = x.shape[-2:]
    if H != W:
        raise ValueError("operands could not be broadcast together: frame (%d, %d) vs distance grid (%d, %d)" % (H, W, W, H))
    lead = tuple(x.shape[:-2])
    frames = 1
    for n in lead:
        frames *= int(n)
    out = torch.empty(max(frames, 1), dtype=torch.float32, device=x.device)
    if frames:
        _lib.check(_lib.lib().fk_electrogram(x.data_ptr(), frames, H, W, float(solve._scalar(point[0])),
                                             float(solve._scalar(point[1])), out.data_ptr(),
                                             ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)))
    return out[:frames].reshape(lead)


def spiral_centres(x):
    """cardiax/metrics.py:25-26 -- a stub in the reference."""
    return None
