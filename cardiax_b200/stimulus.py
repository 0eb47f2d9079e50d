"""Stimulus containers and mask builders -- the reference's API (cardiax/stimulus.py:13-140).

The builders are setup code: masks are made on the host with NumPy/SciPy (the reference also calls
``scipy.ndimage.rotate`` on the host, stimulus.py:137-139) and handed over as fp32 ``torch`` tensors
on the current CUDA device (CPU tensors when no GPU is present, so that setup code can be unit
tested; the solver itself refuses to run without the CUDA library).
"""
from enum import IntEnum
from typing import NamedTuple, Tuple

import numpy as np
import torch

Shape = Tuple[int, ...]
Point2D = Tuple[int, int]


class Protocol(NamedTuple):
    """stimulus.py:13-16 -- start, duration, period in STEP units."""
    start: int
    duration: int
    period: int


class Stimulus(NamedTuple):
    """stimulus.py:19-21."""
    protocol: Protocol
    field: torch.Tensor


class Direction(IntEnum):
    """stimulus.py:24-28."""
    NORTH = 0
    EAST = 1
    SOUTH = 2
    WEST = 3


def _device():
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


def _to_tensor(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(_device())


def _scalar(x):
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    return np.asarray(x).reshape(-1)[0]


def rectangular(shape: Shape, centre: Point2D, size: Point2D, modulus: float, protocol: Protocol):
    """stimulus.py:31-60 -- ``mask[x1:x2, y1:y2] = modulus`` with x on axis 0."""
    mask = np.zeros(shape, dtype=np.float32)
    x1 = int(_scalar(centre[0]) - _scalar(size[0]) / 2)
    x2 = int(_scalar(centre[0]) + _scalar(size[0]) / 2)
    y1 = int(_scalar(centre[1]) - _scalar(size[1]) / 2)
    y2 = int(_scalar(centre[1]) + _scalar(size[1]) / 2)
    mask[x1:x2, y1:y2] = modulus
    return Stimulus(protocol, _to_tensor(mask))


def _linear_np(shape, direction, coverage, modulus):
    stripe_size = int(shape[0] * _scalar(coverage))  # stimulus.py:88 -- shape[0] for every direction
    field = np.zeros(shape, dtype=np.float32)
    direction = _scalar(direction)
    if direction == Direction.WEST:
        field[:, :stripe_size] = modulus
    elif direction == Direction.EAST:
        field[:, -stripe_size:] = modulus
    elif direction == Direction.NORTH:
        field[:stripe_size, :] = modulus
    elif direction == Direction.SOUTH:
        field[-stripe_size:, :] = modulus
    else:
        raise ValueError("direction mus be either 'left', 'right', 'up', or 'down' not %s" % direction)
    return field


def linear(shape: Shape, direction: Direction, coverage: float, modulus: float, protocol: Protocol):
    """stimulus.py:63-106 -- an edge stripe."""
    return Stimulus(protocol, _to_tensor(_linear_np(shape, direction, coverage, modulus)))


def triangular(shape: Shape, direction: float, angle: float, coverage: float, modulus: float, protocol: Protocol):
    """stimulus.py:109-140 -- ``linear`` rotated by ``angle`` degrees (cubic spline, nearest edge mode)."""
    from scipy.ndimage import rotate

    field = _linear_np(shape, direction, coverage, modulus)
    field = rotate(field, angle=float(_scalar(angle)), mode="nearest", prefilter=False, reshape=False)
    return Stimulus(protocol, _to_tensor(field))
