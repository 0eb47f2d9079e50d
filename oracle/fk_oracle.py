"""NumPy restatement of the reference's Fenton-Karma step.  TEST INFRASTRUCTURE ONLY.

Every function follows the reference line by line (paths relative to
``/root/reference``) and keeps its operation ORDER, so that with
``dtype=np.float32`` every intermediate is rounded to fp32 exactly where the
reference's fp32 ``jnp`` arrays are.  ``dtype=np.float64`` gives the twin used
to measure the fp32 drift envelope (the tolerance definition).

What cannot be restated exactly: the reference executes under XLA (jaxlib
0.1.64 pinned in ``install_jax.sh:2``), which may contract ``a*b+c`` to FMA and
supplies its own ``tanh``.  ``tanh="xla"`` restates XLA's published fp32 rational
approximation (``xla/service/llvm_ir/math_ops.cc``, ``EmitFastTanh``: clamp to
[-9, 9], degree-13/degree-6 rational in Horner form, ``|x| < 0.0004 -> x``) with
un-contracted fp32 operations; ``tanh="libm"`` uses ``np.tanh``.  PARITY UNPINNED
for u/v/w values (see ``oracle/__init__.py``).
"""
from typing import NamedTuple, Sequence

import numpy as np


# --------------------------------------------------------------------------- containers
class Params(NamedTuple):
    """cardiax/params.py:4-18 (field ORDER matters)."""
    tau_v_plus: float
    tau_v1_minus: float
    tau_v2_minus: float
    tau_w_plus: float
    tau_w_minus: float
    tau_d: float
    tau_0: float
    tau_r: float
    tau_si: float
    k: float
    V_csi: float
    V_c: float
    V_v: float
    Cm: float


class State(NamedTuple):
    """cardiax/solve.py:12-15 -- (v, w, u), u LAST."""
    v: np.ndarray
    w: np.ndarray
    u: np.ndarray


class Protocol(NamedTuple):
    """cardiax/stimulus.py:13-16 (step units)."""
    start: float
    duration: float
    period: float


class Stimulus(NamedTuple):
    """cardiax/stimulus.py:19-21."""
    protocol: Protocol
    field: np.ndarray


_MAXFLOAT = 1e6  # cardiax/params.py:21
PARAMSETS = {  # cardiax/params.py:23-103
    "1A": Params(3.33, 19.6, 1000, 667, 11, 0.41, 8.3, 50, 45, 10, 0.85, 0.13, 0.0055, 1),
    "1B": Params(3.33, 19.6, 1000, 667, 11, 0.392, 8.3, 50, 45, 10, 0.85, 0.13, 0.0055, 1),
    "1C": Params(3.33, 19.6, 1000, 667, 11, 0.381, 8.3, 50, 45, 10, 0.85, 0.13, 0.0055, 1),
    "1D": Params(3.33, 19.6, 1000, 667, 11, 0.36, 8.3, 50, 45, 10, 0.85, 0.13, 0.0055, 1),
    "1E": Params(3.33, 19.6, 1000, 667, 11, 0.25, 8.3, 50, 45, 10, 0.85, 0.13, 0.0055, 1),
    "2": Params(10, 10, 10, _MAXFLOAT, _MAXFLOAT, 0.25, 10, 190, _MAXFLOAT, 100000, _MAXFLOAT, 0.13, _MAXFLOAT, 1),
    "3": Params(3.33, 19.6, 1250, 870, 41, 0.25, 12.5, 33.33, 29, 10, 0.85, 0.13, 0.04, 1),
    "4A": Params(3.33, 15.6, 5, 350, 80, 0.407, 9, 34, 26.5, 15, 0.45, 0.15, 0.04, 1),
    "4B": Params(3.33, 15.6, 5, 350, 80, 0.405, 9, 34, 26.5, 15, 0.45, 0.15, 0.04, 1),
    "4C": Params(3.33, 15.6, 5, 350, 80, 0.4, 9, 34, 26.5, 15, 0.45, 0.15, 0.04, 1),
    "5": Params(3.33, 12, 2, 1000, 100, 0.362, 5, 33.33, 29, 15, 0.7, 0.13, 0.04, 1),
    "6": Params(3.33, 9, 8, 250, 60, 0.395, 9, 33.33, 29, 15, 0.5, 0.13, 0.04, 1),
    "7": Params(10, 7, 7, _MAXFLOAT, _MAXFLOAT, 0.25, 12, 100, _MAXFLOAT, _MAXFLOAT, _MAXFLOAT, 0.13, _MAXFLOAT, 1),
    "8": Params(13.03, 19.06, 1250, 800, 40, 0.45, 12.5, 33.25, 29, 10, 0.85, 0.13, 0.04, 1),
    "9": Params(3.33, 15, 2, 670, 61, 0.25, 12.5, 28, 29, 10, 0.45, 0.13, 0.05, 1),
    "10": Params(10, 40, 333, 1000, 65, 0.115, 12.5, 25, 22.22, 10, 0.85, 0.13, 0.0025, 1),
}


# --------------------------------------------------------------------------- tanh
_TANH_NUM = [np.float32(c) for c in (
    -2.76076847742355e-16, 2.00018790482477e-13, -8.60467152213735e-11,
    5.12229709037114e-08, 1.48572235717979e-05, 6.37261928875436e-04,
    4.89352455891786e-03)]
_TANH_DEN = [np.float32(c) for c in (
    1.19825839466702e-06, 1.18534705686654e-04, 2.26843463243900e-03,
    4.89352518554385e-03)]


def tanh_xla_f32(x):
    """XLA's fp32 tanh (third-party: jaxlib 0.1.64, EmitFastTanh), un-contracted fp32.

    clamp to [-9, 9]; x2 = x*x; num = x*Horner(x2, 7 coeffs); den = Horner(x2, 4 coeffs);
    result = |x| < 0.0004 ? x : num/den.
    """
    x = np.asarray(x, dtype=np.float32)
    xc = np.minimum(np.maximum(x, np.float32(-9.0)), np.float32(9.0))
    x2 = xc * xc
    num = np.full_like(x2, _TANH_NUM[0])
    for c in _TANH_NUM[1:]:
        num = x2 * num + c
    num = xc * num
    den = np.full_like(x2, _TANH_DEN[0])
    for c in _TANH_DEN[1:]:
        den = x2 * den + c
    return np.where(np.abs(x) < np.float32(0.0004), x, num / den)


# --------------------------------------------------------------------------- solve.py
def init(shape, dtype=np.float32):
    """cardiax/solve.py:18-23 -- v = 1, w = 1, u = 0."""
    return State(np.ones(shape, dtype), np.ones(shape, dtype), np.zeros(shape, dtype))


def gradient(a, axis):
    """cardiax/solve.py:225-254 -- first derivative times dx; 3rd-order edges, 4th-order inner.

    Evaluated left to right as written; coefficients are Python doubles that take the
    array dtype (weak typing), each product rounded before the next add.
    """
    a = np.asarray(a)
    dt = a.dtype.type
    n = a.shape[axis]

    def sl(lo, hi):  # jax.lax.slice_in_dim(a, lo, hi, axis=axis) with Python slice semantics
        idx = [slice(None)] * a.ndim
        idx[axis] = slice(lo, hi)
        return a[tuple(idx)]

    assert n >= 5
    lo = (dt(-11 / 6) * sl(0, 2) + dt(3) * sl(1, 3) - dt(3 / 2) * sl(2, 4) + dt(1 / 3) * sl(3, 5))
    mid = (dt(1 / 12) * sl(None, -4) - dt(2 / 3) * sl(1, -3) + dt(2 / 3) * sl(3, -1) - dt(1 / 12) * sl(4, None))
    hi = (dt(-1 / 3) * sl(-5, -3) + dt(3 / 2) * sl(-4, -2) - dt(3) * sl(-3, -1) + dt(11 / 6) * sl(-2, None))
    return np.concatenate((lo, mid, hi), axis)


def stimulus_active(t, protocol, dtype=np.float32):
    """cardiax/solve.py:262-267 -- the scalar predicate, in the counter's dtype."""
    t = dtype(t)
    start = dtype(np.asarray(protocol.start).reshape(-1)[0])
    duration = dtype(np.asarray(protocol.duration).reshape(-1)[0])
    period = dtype(np.asarray(protocol.period).reshape(-1)[0])
    active = t >= start
    with np.errstate(invalid="ignore"):
        active &= np.mod(start - t + dtype(1), period) < duration
    return bool(active)


def stimulus_active_typed(t, protocol):
    """cardiax/solve.py:262-267 with jax's typing (x64 disabled): the counter and each protocol entry is an int32 or a
    float32 according to what the caller passed -- Python int / integer array -> int32 (``deepx/generate.py:24-27``
    draws ``start`` and ``period`` as shape-(1,) int32 arrays and ``generate.sequence`` :187-196 runs an int32 counter),
    Python float / float array -> float32 (``solve.forward`` :210-211 passes ``float(checkpoint)``) -- and every binary
    operation runs in float32 as soon as one side is a float, in int32 otherwise."""
    def canon(x):
        a = np.asarray(x).reshape(-1)[0]
        return np.int32(a) if a.dtype.kind in "iub" else np.float32(a)

    def both(a, b):
        if isinstance(a, np.float32) or isinstance(b, np.float32):
            return np.float32(a), np.float32(b)
        return a, b

    t, start, duration, period = canon(t), canon(protocol.start), canon(protocol.duration), canon(protocol.period)
    a, b = both(t, start)
    active = bool(a >= b)
    a, b = both(start, t)
    with np.errstate(all="ignore"):
        x = a - b
        x = x + type(x)(1)
        a, b = both(x, period)
        m = np.mod(a, b)                      # jnp.mod: sign of the divisor, for ints and floats alike
    a, b = both(m, duration)
    return active and bool(a < b)


def stimulate(t, X, stimuli: Sequence[Stimulus], dtype=np.float32, typed=False):
    """cardiax/solve.py:257-271 -- later stimuli override earlier; zero cells never stimulate.  ``typed``: evaluate the
    schedule with the callers' own int / float types (`stimulus_active_typed`) instead of casting everything to ``dtype``."""
    X = np.asarray(X)
    stimulated = np.zeros_like(X)
    for s in stimuli:
        active = stimulus_active_typed(t, s.protocol) if typed else stimulus_active(t, s.protocol, dtype)
        field = np.asarray(s.field, dtype=X.dtype)
        stimulated = np.where((field * X.dtype.type(active)) != 0, field, stimulated)
    return np.where(stimulated != 0, stimulated, X)


def step(state: State, t, params: Params, diffusivity, stimuli, dx, dtype=np.float32, tanh="xla", typed=False):
    """cardiax/solve.py:26-65 -- RHS (d_v, d_w, d_u) at step index t."""
    f = dtype
    P = Params(*[f(x) for x in params])
    dx = f(dx)
    one = f(1)
    # :29-32 neumann boundary conditions
    v = np.pad(np.asarray(state.v, f), 1, mode="edge")
    w = np.pad(np.asarray(state.w, f), 1, mode="edge")
    u = np.pad(np.asarray(state.u, f), 1, mode="edge")
    D = np.pad(np.asarray(diffusivity, f), 1, mode="edge")

    # :35-37 reaction term
    p = (u >= P.V_c).astype(f)
    q = (u >= P.V_v).astype(f)
    tau_v_minus = (one - q) * P.tau_v1_minus + q * P.tau_v2_minus

    # :39-42
    j_fi = -v * p * (u - P.V_c) * (one - u) / P.tau_d
    j_so = (u * (one - p) / P.tau_0) + (p / P.tau_r)
    arg = P.k * (u - P.V_csi)
    if f is np.float32 and tanh == "xla":
        th = tanh_xla_f32(arg)
    else:
        th = np.tanh(arg)
    j_si = -(w * (one + th)) / (f(2) * P.tau_si)
    j_ion = -(j_fi + j_so + j_si) / P.Cm

    # :45-46 stimulus REPLACES j_ion
    stimuli = [Stimulus(s.protocol, np.pad(np.asarray(s.field, f), 1, mode="edge")) for s in stimuli]
    j_ion = stimulate(t, j_ion, stimuli, dtype=f, typed=typed)

    # :49-55 diffusion term
    u_x = gradient(u, 0) / dx
    u_y = gradient(u, 1) / dx
    u_xx = gradient(u_x, 0) / dx
    u_yy = gradient(u_y, 1) / dx
    D_x = gradient(D, 0) / dx
    D_y = gradient(D, 1) / dx
    del_u = D * (u_xx + u_yy) + (D_x * u_x) + (D_y * u_y)

    # :57-59
    d_v = ((one - p) * (one - v) / tau_v_minus) - ((p * v) / P.tau_v_plus)
    d_w = ((one - p) * (one - w) / P.tau_w_minus) - ((p * w) / P.tau_w_plus)
    d_u = del_u + j_ion
    # :61-65
    return State(d_v[1:-1, 1:-1], d_w[1:-1, 1:-1], d_u[1:-1, 1:-1])


def step_euler(state, t, params, diffusivity, stimuli, dt, dx, dtype=np.float32, tanh="xla", typed=False):
    """cardiax/solve.py:68-70 -- x + d_x*dt."""
    g = step(state, t, params, diffusivity, stimuli, dx, dtype=dtype, tanh=tanh, typed=typed)
    dt = dtype(dt)
    return State(*[np.add(np.asarray(x, dtype), dx_ * dt) for x, dx_ in zip(state, g)])


def _counter(t, t_end, counter):
    """The values ``lax.fori_loop(t, t_end, ...)`` hands to the body: ``i = t; while i < t_end: yield i; i = i + 1`` in
    the bounds' dtype.  "f32": the float32 counter of ``solve.forward``; "i32": the int32 counter of
    ``deepx.generate.sequence``; None: a Python float (exact below 2^53; identical to "f32" below 2^24)."""
    if counter == "f32":
        i, end, one = np.float32(t), np.float32(t_end), np.float32(1)
    elif counter == "i32":
        i, end, one = np.int32(t), np.int32(t_end), np.int32(1)
    else:
        i, end, one = float(t), float(t_end), 1.0
    while i < end:
        yield i
        nxt = i + one
        if nxt == i:
            raise OverflowError("the reference's float32 loop counter stops advancing at 2^24")
        i = nxt


def forward_euler(state, t, t_end, params, diffusivity, stimuli, dt, dx, dtype=np.float32, tanh="xla", counter=None):
    """cardiax/solve.py:92-100 -- fori_loop over [t, t_end) with a counter of t's dtype (see `_counter`)."""
    for i in _counter(t, t_end, counter):
        state = step_euler(state, i, params, diffusivity, stimuli, dt, dx, dtype=dtype, tanh=tanh,
                           typed=counter is not None)
    return state


def step_heun(state, t, params, diffusivity, stimuli, dt, dx, dtype=np.float32, tanh="xla", typed=False):
    """cardiax/solve.py:73-85 -- Heun: k1 at y, k2 at y + k1*dt (SAME counter t), y + (k1 + k2) * (dt * 0.5)."""
    def euler(y, dy, h):  # :74-75  jnp.add(v, dv * h)
        return State(*[np.add(np.asarray(a, dtype), b * h) for a, b in zip(y, dy)])
    d_state = step(state, t, params, diffusivity, stimuli, dx, dtype=dtype, tanh=tanh, typed=typed)
    new_state = euler(state, d_state, dtype(dt))
    d_new_state = step(new_state, t, params, diffusivity, stimuli, dx, dtype=dtype, tanh=tanh, typed=typed)
    summed = State(*[np.add(a, b) for a, b in zip(d_state, d_new_state)])
    return euler(state, summed, dtype(dt) * dtype(0.5))   # `dt * 0.5`: dt is a 32-bit scalar inside the jitted loop


def forward_heun(state, t, t_end, params, diffusivity, stimuli, dt, dx, dtype=np.float32, tanh="xla", counter=None):
    """cardiax/solve.py:103-111."""
    for i in _counter(t, t_end, counter):
        state = step_heun(state, i, params, diffusivity, stimuli, dt, dx, dtype=dtype, tanh=tanh,
                          typed=counter is not None)
    return state


# --------------------------------------------------------------------------- stimulus.py
def rectangular(shape, centre, size, modulus, protocol):
    """cardiax/stimulus.py:31-60."""
    mask = np.zeros(shape, dtype=np.float32)
    x1 = int(centre[0] - size[0] / 2)
    x2 = int(centre[0] + size[0] / 2)
    y1 = int(centre[1] - size[1] / 2)
    y2 = int(centre[1] + size[1] / 2)
    mask[x1:x2, y1:y2] = modulus
    return Stimulus(protocol, mask)


def linear(shape, direction, coverage, modulus, protocol):
    """cardiax/stimulus.py:63-106 -- NORTH=0 rows[:s], EAST=1 cols[-s:], SOUTH=2 rows[-s:], WEST=3 cols[:s]."""
    stripe = int(shape[0] * coverage)
    field = np.zeros(shape, dtype=np.float32)
    d = int(direction)
    if d == 3:
        field[:, :stripe] = modulus
    elif d == 1:
        field[:, -stripe:] = modulus
    elif d == 0:
        field[:stripe, :] = modulus
    elif d == 2:
        field[-stripe:, :] = modulus
    else:
        raise ValueError("direction mus be either 'left', 'right', 'up', or 'down' not %s" % direction)
    return Stimulus(protocol, field)


def triangular(shape, direction, angle, coverage, modulus, protocol):
    """cardiax/stimulus.py:109-140."""
    from scipy.ndimage import rotate
    s = linear(shape, direction, coverage, modulus, protocol)
    return Stimulus(protocol, rotate(s.field, angle=angle, mode="nearest", prefilter=False, reshape=False))
