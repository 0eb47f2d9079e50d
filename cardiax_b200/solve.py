"""The reference's solver API (cardiax/solve.py) on top of libfk.so.

Same names, positional order and defaults as the reference; arrays are fp32 ``torch`` CUDA tensors
instead of ``jax`` device arrays.  Inputs are never mutated, every call returns new tensors and is
enqueued on the current torch CUDA stream without synchronising.

Extension (not in the reference): every function that takes a ``State`` also accepts a batch of
independent tissues, ``(batch, H, W)`` tensors, with ``diffusivity`` either ``(H, W)`` or
``(batch, H, W)`` and ``stimuli`` either one list shared by all tissues or one list per tissue.

Limit (not in the reference): at most 32 stimuli per tissue (the kernels carry the set of active stimuli of a level as
a 32-bit mask); a longer list raises ``ValueError``.  The reference's own drivers use 1-3.
"""
import ctypes
from enum import Enum
from typing import NamedTuple

import numpy as np
import torch

from . import _lib, convert, options


class State(NamedTuple):
    """cardiax/solve.py:12-15 -- (v, w, u): u is LAST."""
    v: torch.Tensor
    w: torch.Tensor
    u: torch.Tensor


# --------------------------------------------------------------------------- helpers
def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("cardiax_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _as_f32(x, device=None):
    device = device or _device()
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(np.asarray(x))
    return x.to(device=device, dtype=torch.float32).contiguous()


def _scalar(x):
    """Python/NumPy/torch scalar or shape-(1,) array -> float (protocol entries, t, dt, dx)."""
    if isinstance(x, torch.Tensor):
        return float(x.detach().reshape(-1)[0].item())
    return float(np.asarray(x).reshape(-1)[0])


def _is_int(x):
    """Does jax (x64 disabled) see this scalar / shape-(1,) array as an int32 rather than a float32?  Python ints, NumPy
    and torch integers and booleans do; everything else is a float.  Decides the typing of the stimulus schedule
    (cardiax/solve.py:262-267): see include/fk.h, FkStimulus::int_mask and FkOptions::counter_is_int."""
    if isinstance(x, torch.Tensor):
        return not (x.dtype.is_floating_point or x.dtype.is_complex)
    if isinstance(x, (bool, int, np.integer, np.bool_)):
        return True
    if isinstance(x, (float, np.floating)):
        return False
    return np.asarray(x).dtype.kind in "iub"


def _stimulus_struct(field_ptr, protocol):
    mask = (1 if _is_int(protocol.start) else 0) | (2 if _is_int(protocol.duration) else 0) | (4 if _is_int(protocol.period) else 0)
    return _lib.FkStimulus(field_ptr, _scalar(protocol.start), _scalar(protocol.duration), _scalar(protocol.period), mask, 0)


def _params_struct(params):
    return _lib.FkParams(*[np.float32(_scalar(p)) for p in params])


def _is_stimulus(s):
    return hasattr(s, "protocol") and hasattr(s, "field")


def _pack_stimuli(stimuli, batch, shape, device):
    """-> (ctypes array of batch*n_stim FkStimulus, n_stim, keep-alive list)."""
    stimuli = list(stimuli)
    if len(stimuli) and not _is_stimulus(stimuli[0]):
        per = [list(s) for s in stimuli]  # one list per tissue
        if len(per) != batch:
            raise ValueError("expected one stimulus list per tissue (%d), got %d" % (batch, len(per)))
    else:
        per = [stimuli] * batch
    n_stim = max(len(p) for p in per) if per else 0
    arr = (_lib.FkStimulus * max(1, batch * n_stim))()
    keep = []
    for b in range(batch):
        for i in range(n_stim):
            if i < len(per[b]):
                s = per[b][i]
                f = _as_f32(s.field, device)
                if tuple(f.shape) != tuple(shape):
                    raise ValueError("stimulus field shape %s != tissue shape %s" % (tuple(f.shape), tuple(shape)))
                keep.append(f)
                arr[b * n_stim + i] = _stimulus_struct(f.data_ptr(), s.protocol)
            else:
                arr[b * n_stim + i] = _lib.FkStimulus(None, 0.0, 0.0, 1.0, 0, 0)
    return arr, n_stim, keep


_uniform_cache = {}   # id(tensor) -> (weakref to that tensor, its _version, verdict)


def _is_uniform(D):
    """Is the diffusivity map one constant?  One device reduction + host read.

    The verdict is remembered only for the caller's OWN tensor object: the entry holds a weak reference and is valid
    while ``ref() is D`` and ``D._version`` is unchanged.  A device address does not identify a tensor -- the caching
    allocator hands a freed block straight back, so a scar map uploaded where a constant map just lived would have
    inherited its verdict -- and tensors this module had to create itself (NumPy / CPU / non-fp32 inputs) are new
    objects on every call, hence always re-examined."""
    import weakref
    hit = _uniform_cache.get(id(D))
    if hit is not None and hit[0]() is D and hit[1] == D._version:
        return hit[2]
    verdict = bool((D.min() == D.max()).item())
    if len(_uniform_cache) > 64:
        for k in [k for k, v in _uniform_cache.items() if v[0]() is None]:
            del _uniform_cache[k]
        if len(_uniform_cache) > 64:
            _uniform_cache.clear()
    _uniform_cache[id(D)] = (weakref.ref(D), D._version, verdict)
    return verdict


_division_checked = {}


def _division_is_safe(P, dx):
    """Exact numerics: verify the FMA division by this run's constants on the device, once per (params, dx)."""
    key = (bytes(P), float(dx))
    hit = _division_checked.get(key)
    if hit is None:
        bad = ctypes.c_longlong(-1)
        _lib.check(_lib.lib().fk_check_exact_division(ctypes.byref(P), np.float32(dx), ctypes.byref(bad), _stream()))
        hit = bad.value == 0
        _division_checked[key] = hit
    return hit


def _options(D=None, P=None, dx=None, **over):
    o = _lib.FkOptions()
    _lib.lib().fk_default_options(ctypes.byref(o))
    o.exact = int(options.numerics == "exact")
    if o.exact and P is not None:
        o.safe_division = int(options.safe_division or not _division_is_safe(P, _scalar(dx)))
    o.steps_per_launch = int(options.steps_per_launch)
    o.kernel = int(options.kernel)
    o.cta_threads = int(options.cta_threads)
    o.rows_per_cta = int(options.rows_per_cta)
    o.tiles_r, o.tiles_c = int(options.tiles[0]), int(options.tiles[1])
    o.cells_per_thread = int(options.cells_per_thread)
    o.edge_rows, o.edge_colgroups = int(options.edge_tile[0]), int(options.edge_tile[1])
    o.maps_global = int(options.maps_global)
    if D is not None and options.detect_uniform_diffusivity:
        o.uniform_diffusivity = int(_is_uniform(D))
    for k, v in over.items():
        setattr(o, k, int(v))
    return o


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _prep(state, diffusivity):
    dev = _device()
    v, w, u = [_as_f32(x, dev) for x in state]
    if not (v.shape == w.shape == u.shape) or u.dim() not in (2, 3):
        raise ValueError("state arrays must share one (H, W) or (batch, H, W) shape")
    batch = u.shape[0] if u.dim() == 3 else 1
    H, W = u.shape[-2:]
    D = _as_f32(diffusivity, dev)
    if tuple(D.shape[-2:]) != (H, W) or D.dim() not in (2, 3) or (D.dim() == 3 and D.shape[0] != batch):
        raise ValueError("diffusivity shape %s does not match the tissue %s" % (tuple(D.shape), tuple(u.shape)))
    return dev, v, w, u, D, batch, H, W


def _workspace(L, H, W, batch, n_stim, d_batched, dev):
    nbytes = L.fk_workspace_bytes(H, W, batch, n_stim, d_batched)
    return torch.empty(nbytes, dtype=torch.uint8, device=dev), nbytes


# --------------------------------------------------------------------------- reference API
def init(shape):
    """cardiax/solve.py:18-23 -- v = 1, w = 1, u = 0."""
    dev = _device()
    shape = tuple(int(s) for s in shape)
    return State(torch.ones(shape, dtype=torch.float32, device=dev), torch.ones(shape, dtype=torch.float32, device=dev),
                 torch.zeros(shape, dtype=torch.float32, device=dev))


def step(state, t, params, diffusivity, stimuli, dx):
    """cardiax/solve.py:26-65 -- time derivatives State(d_v, d_w, d_u) at counter ``t``."""
    L = _lib.lib()
    dev, v, w, u, D, batch, H, W = _prep(state, diffusivity)
    arr, n_stim, keep = _pack_stimuli(stimuli, batch, (H, W), dev)
    ws, nbytes = _workspace(L, H, W, batch, n_stim, int(D.dim() == 3), dev)
    dv, dw, du = torch.empty_like(v), torch.empty_like(w), torch.empty_like(u)
    P = _params_struct(params)
    o = _options(None, P, dx, counter_is_int=_is_int(t))
    _lib.check(L.fk_rhs(v.data_ptr(), w.data_ptr(), u.data_ptr(), dv.data_ptr(), dw.data_ptr(), du.data_ptr(),
                        D.data_ptr(), int(D.dim() == 3), H, W, batch, ctypes.byref(P), arr, n_stim, _scalar(t),
                        np.float32(_scalar(dx)), ctypes.byref(o), ws.data_ptr(), nbytes, _stream()))
    return State(dv, dw, du)


def _forward_euler(state, t, t_end, params, diffusivity, stimuli, dt, dx):
    """cardiax/solve.py:92-100 -- ``lax.fori_loop(t, t_end, step_euler)``: Euler steps for counter in [t, t_end)."""
    L = _lib.lib()
    dev, v, w, u, D, batch, H, W = _prep(state, diffusivity)
    arr, n_stim, keep = _pack_stimuli(stimuli, batch, (H, W), dev)
    ws, nbytes = _workspace(L, H, W, batch, n_stim, int(D.dim() == 3), dev)
    vo, wo, uo = torch.empty_like(v), torch.empty_like(w), torch.empty_like(u)
    P = _params_struct(params)
    o = _options(D, P, dx, counter_is_int=_is_int(t) and _is_int(t_end))   # the fori_loop counter takes the bounds' dtype
    _lib.check(L.fk_forward_euler(v.data_ptr(), w.data_ptr(), u.data_ptr(), vo.data_ptr(), wo.data_ptr(), uo.data_ptr(),
                                  D.data_ptr(), int(D.dim() == 3), H, W, batch, ctypes.byref(P), arr, n_stim,
                                  _scalar(t), _scalar(t_end), np.float32(_scalar(dt)), np.float32(_scalar(dx)),
                                  ctypes.byref(o), ws.data_ptr(), nbytes, _stream()))
    return State(vo, wo, uo)


def step_euler(state, t, params, diffusivity, stimuli, dt, dx):
    """cardiax/solve.py:68-70 -- one Euler step ``x + d_x * dt``."""
    t1 = (int(_scalar(t)) + 1) if _is_int(t) else (_scalar(t) + 1.0)
    return _forward_euler(state, t, t1, params, diffusivity, stimuli, dt, dx)


def step_heun(state, t, params, diffusivity, stimuli, dt, dx):
    """cardiax/solve.py:73-85 -- Heun: both stages evaluated at the same counter ``t``."""
    t1 = (int(_scalar(t)) + 1) if _is_int(t) else (_scalar(t) + 1.0)
    return _forward_heun(state, t, t1, params, diffusivity, stimuli, dt, dx)


def _forward_heun(state, t, t_end, params, diffusivity, stimuli, dt, dx):
    """cardiax/solve.py:103-111 -- ``lax.fori_loop(t, t_end, step_heun)``; the whole loop runs on the device."""
    L = _lib.lib()
    dev, v, w, u, D, batch, H, W = _prep(state, diffusivity)
    arr, n_stim, keep = _pack_stimuli(stimuli, batch, (H, W), dev)
    nbytes = L.fk_heun_workspace_bytes(H, W, batch, n_stim, int(D.dim() == 3))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    vo, wo, uo = torch.empty_like(v), torch.empty_like(w), torch.empty_like(u)
    P = _params_struct(params)
    o = _options(D, P, dx, counter_is_int=_is_int(t) and _is_int(t_end))
    _lib.check(L.fk_forward_heun(v.data_ptr(), w.data_ptr(), u.data_ptr(), vo.data_ptr(), wo.data_ptr(), uo.data_ptr(),
                                 D.data_ptr(), int(D.dim() == 3), H, W, batch, ctypes.byref(P), arr, n_stim,
                                 _scalar(t), _scalar(t_end), np.float32(_scalar(dt)), np.float32(_scalar(dx)),
                                 ctypes.byref(o), ws.data_ptr(), nbytes, _stream()))
    return State(vo, wo, uo)


def _forward_dormandprince(state, ts, params, diffusivity, stimuli, dt, dx):
    """cardiax/solve.py:114-124 -- ``jax.experimental.ode.odeint(step, state, ts, params, diffusivity, stimuli, dx)``:
    adaptive Dormand-Prince with jax's controller and dense output; ``ts`` are continuous times in the same (step)
    units the checkpoints are given in and ``dt`` is unused, exactly like the reference.  Returns a State of stacked
    arrays ``(len(ts), H, W)`` whose first entry is the initial state.  The arrays stay on the device; the step-size
    controller runs on the host, so this call synchronises the current stream (options.ode_* = odeint's keywords)."""
    L = _lib.lib()
    dev, v, w, u, D, batch, H, W = _prep(state, diffusivity)
    arr, n_stim, keep = _pack_stimuli(stimuli, batch, (H, W), dev)
    if isinstance(ts, torch.Tensor):
        ts = ts.detach().cpu().numpy()
    ts = np.ascontiguousarray(np.asarray(ts, dtype=np.float64).reshape(-1), dtype=np.float32)
    n_ts = int(ts.shape[0])
    nbytes = L.fk_dopri5_workspace_bytes(H, W, batch, n_stim, int(D.dim() == 3))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    vo, wo, uo = [torch.empty((n_ts,) + tuple(u.shape), dtype=torch.float32, device=dev) for _ in range(3)]
    P = _params_struct(params)
    o = _options(None, P, dx)
    stats = (ctypes.c_longlong * 3)()
    _lib.check(L.fk_odeint_dopri5(v.data_ptr(), w.data_ptr(), u.data_ptr(), vo.data_ptr(), wo.data_ptr(), uo.data_ptr(),
                                  D.data_ptr(), int(D.dim() == 3), H, W, batch, ctypes.byref(P), arr, n_stim,
                                  ts.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), n_ts, np.float32(_scalar(dx)),
                                  np.float32(options.ode_rtol), np.float32(options.ode_atol), float(options.ode_mxstep),
                                  ctypes.byref(o), ws.data_ptr(), nbytes, _stream(), ctypes.byref(stats)))
    global last_ode_stats
    last_ode_stats = dict(attempts=stats[0], accepted=stats[1], rhs_evals=stats[2])
    return State(vo, wo, uo)


last_ode_stats = None


def step_rk(state, t, params, diffusivity, stimuli, dt, dx):
    """cardiax/solve.py:88-89 -- ``ode.odeint(step, state, t, ...)``: ``t`` is the array of output times."""
    return _forward_dormandprince(state, t, params, diffusivity, stimuli, dt, dx)


class TimeIntegrator(Enum):
    """cardiax/solve.py:127-130 -- as in the reference the attributes ARE the forward functions."""
    EULER = _forward_euler
    HEUN = _forward_heun
    DORMANDPRINCE = _forward_dormandprince


def forward_dimensional(tissue_size, final_time, ms_step, params, diffusivity, stimuli, dt, dx,
                        integrator=TimeIntegrator.EULER, plot_while=False):
    """cardiax/solve.py:133-165 -- physical units (cm, ms) wrapper around ``forward``."""
    shape = convert.realsize_to_shape(tissue_size, dx)
    start = 0
    stop = convert.ms_to_units(final_time, dt)
    step_ = convert.ms_to_units(ms_step, dt)

    assert shape == tuple(diffusivity.shape)
    assert all([tuple(s.field.shape) == shape for s in stimuli])

    state = init(shape)
    checkpoints = np.arange(start, stop, step_)
    return forward(state, checkpoints, params, diffusivity, stimuli, dt, dx, integrator, plot_while)


def forward(state, checkpoints, params, diffusivity, stimuli, dt, dx, integrator=TimeIntegrator.EULER,
            plot_while=False):
    """cardiax/solve.py:168-222 -- checkpoint loop; returns the list of States at checkpoints[1:]."""
    if plot_while:
        from . import plot  # matplotlib is optional
        plot.plot_stimuli(stimuli)
        plot.plot_diffusivity(diffusivity)
        plot.plot_state(state)

    f = integrator
    if f == TimeIntegrator.DORMANDPRINCE:
        return f(state, checkpoints, params, diffusivity, stimuli, dt, dx)

    cps = [_scalar(c) for c in checkpoints]
    dtf = _scalar(dt)
    states = []
    for i in range(len(cps) - 1):
        if options.verbose:
            print("Solving at: %dms/%dms\t\t with %d passages" % (cps[i + 1] * dtf, cps[-1] * dtf, cps[i + 1] - cps[i]),
                  end="\r")
        state = f(state, float(cps[i]), float(cps[i + 1]), params, diffusivity, stimuli, dt, dx)
        if plot_while:
            from . import plot
            plot.plot_state(state)
        states.append(state)
    return states


def gradient(a, axis):
    """cardiax/solve.py:225-254 -- first derivative times dx along ``axis`` of an N-D array (n >= 5)."""
    L = _lib.lib()
    a = _as_f32(a)
    axis = int(axis)
    if axis < 0:
        axis += a.dim()
    n = a.shape[axis]
    outer = int(np.prod(a.shape[:axis], dtype=np.int64)) if axis > 0 else 1
    inner = int(np.prod(a.shape[axis + 1:], dtype=np.int64)) if axis + 1 < a.dim() else 1
    out = torch.empty_like(a)
    _lib.check(L.fk_gradient(a.data_ptr(), out.data_ptr(), outer, n, inner, _stream()))
    return out


def stimulate(t, X, stimuli):
    """cardiax/solve.py:257-271 -- X with the active stimuli written over it (later stimuli win)."""
    L = _lib.lib()
    X = _as_f32(X)
    if X.dim() != 2:
        raise ValueError("stimulate expects a 2-D array")
    H, W = X.shape
    arr, n_stim, keep = _pack_stimuli(stimuli, 1, (H, W), X.device)
    ws = torch.empty(max(64, 64 * n_stim), dtype=torch.uint8, device=X.device)
    out = torch.empty_like(X)
    _lib.check(L.fk_stimulate(_scalar(t), int(_is_int(t)), X.data_ptr(), out.data_ptr(), H, W, arr, n_stim, ws.data_ptr(), ws.numel(),
                              _stream()))
    return out
