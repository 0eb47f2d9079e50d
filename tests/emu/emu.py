"""ctypes wrapper of tests/emu/fk_emu.cpp -- CPU emulation of the CUDA kernel bodies.  TESTS ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfk_emu.so")
_CSRC = os.path.join(_HERE, "..", "..", "cardiax_b200", "csrc")
_lib = None


class _Stim(ctypes.Structure):
    _fields_ = [("field", ctypes.c_void_p), ("start", ctypes.c_double), ("duration", ctypes.c_double),
                ("period", ctypes.c_double), ("int_mask", ctypes.c_int), ("reserved", ctypes.c_int)]


def _is_int(x):
    return np.asarray(x).dtype.kind in "iub" and not isinstance(x, float)


def build(force=False):
    deps = [os.path.join(_HERE, "fk_emu.cpp")] + [os.path.join(_CSRC, f) for f in
                                                    ("fk_core.h", "fk_tile.h", "fk_stream.h", "fk_driver.h", "fk_wide.h", "fk_resident.h", "fk_aux.h", "fk_ode.h")]
    def fresh():
        return os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(d) for d in deps)
    if not force and fresh():
        return _SO
    import fcntl   # (the two processes of a gloo test may both find the library stale: one builds, the link is renamed into place)
    with open(_SO + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if force or not fresh():
                tmp = _SO + ".tmp.%d" % os.getpid()
                subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
                                       "-Wno-unknown-pragmas", "-o", tmp, deps[0]])
                os.replace(tmp, _SO)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.fk_emu_euler.restype = ctypes.c_int
        _lib.fk_emu_stim_active.restype = ctypes.c_int
        _lib.fk_emu_stim_active.argtypes = [ctypes.c_float] * 4
    return _lib


def stim_active(t, start, duration, period):
    return bool(lib().fk_emu_stim_active(t, start, duration, period))


def stim_active_typed(t, protocol):
    """fk::stim_active_typed with the typing of the Python objects handed in (the oracle's stimulus_active_typed)."""
    L = lib()
    L.fk_emu_stim_active_typed.restype = ctypes.c_int
    L.fk_emu_stim_active_typed.argtypes = [ctypes.c_double, ctypes.c_int] + [ctypes.c_double] * 3 + [ctypes.c_int]
    mask = sum(bit for bit, x in zip((1, 2, 4), protocol) if _is_int(x))
    return bool(L.fk_emu_stim_active_typed(float(np.asarray(t).reshape(-1)[0]), int(_is_int(t)),
                                           *[float(np.asarray(x).reshape(-1)[0]) for x in protocol], mask))


def stims_quiet(t0, nsteps, protocol):
    """fk::stims_quiet for one stimulus with the typing of the Python objects handed in."""
    L = lib()
    L.fk_emu_stims_quiet.restype = ctypes.c_int
    L.fk_emu_stims_quiet.argtypes = [ctypes.c_double, ctypes.c_longlong] + [ctypes.c_double] * 3 + [ctypes.c_int]
    mask = sum(bit for bit, x in zip((1, 2, 4), protocol) if _is_int(x))
    return bool(L.fk_emu_stims_quiet(float(t0), int(nsteps), *[float(np.asarray(x).reshape(-1)[0]) for x in protocol], mask))


def euler(state, t0, t1, params, D, stimuli, dt, dx, exact=True, T=0, kernel=0, cta_threads=0, rows_per_cta=0,
          uniform=0, reverse=0, phys_top=1, phys_bottom=1, rhs=False, row0=0, row1=0, tiles=(0, 0), nc=0, edge_tile=(0, 0), maps_global=0,
          typed=False):
    """state: (v, w, u) arrays of shape (H, W) or (batch, H, W).  Returns (v, w, u) and launch counts.
    typed: keep the int / float typing of t0, t1 and the protocol entries (the reference's int32-counter path) instead
    of the all-float typing of `solve.forward`."""
    v, w, u = [np.ascontiguousarray(x, dtype=np.float32) for x in state]
    batched = u.ndim == 3
    batch = u.shape[0] if batched else 1
    H, W = u.shape[-2:]
    D = np.ascontiguousarray(D, dtype=np.float32)
    d_batched = int(D.ndim == 3)
    vo, wo, uo = np.full_like(v, np.nan), np.full_like(w, np.nan), np.full_like(u, np.nan)
    par = np.array([float(np.asarray(x).reshape(-1)[0]) for x in params], dtype=np.float32)
    # stimuli: list (shared by every tissue) or list of lists (per tissue)
    per = stimuli if (len(stimuli) and isinstance(stimuli[0], (list, tuple)) and not hasattr(stimuli[0], "protocol")) \
        else [stimuli] * batch
    n_stim = len(per[0])
    keep = []
    arr = (_Stim * max(1, batch * n_stim))()
    for b in range(batch):
        for i, s in enumerate(per[b]):
            f = np.ascontiguousarray(s.field, dtype=np.float32)
            keep.append(f)
            mask = sum(bit for bit, x in zip((1, 2, 4), s.protocol) if typed and _is_int(x))
            arr[b * n_stim + i] = _Stim(f.ctypes.data, *[float(np.asarray(x).reshape(-1)[0]) for x in s.protocol], mask, 0)
    opts = (ctypes.c_int * 18)(int(exact), T, kernel, phys_top, phys_bottom, cta_threads, rows_per_cta, uniform, reverse,
                               row0, row1, int(tiles[0]), int(tiles[1]), int(nc), int(edge_tile[0]), int(edge_tile[1]),
                               int(maps_global), int(typed and _is_int(t0) and _is_int(t1)))
    info = (ctypes.c_int * 2)()
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib().fk_emu_euler(p(v), p(w), p(u), p(vo), p(wo), p(uo), p(D), d_batched, H, W, batch, p(par), arr, n_stim,
                            ctypes.c_double(t0), ctypes.c_double(t1), ctypes.c_float(dt), ctypes.c_float(dx), opts,
                            int(rhs), info)
    if rc != 0:
        raise RuntimeError("fk_emu_euler rc=%d" % rc)
    return (vo, wo, uo), (info[0], info[1])


def set_mirror(up, down, row0, row1, dst_row0):
    """Halo mirror of the next `euler` call: up / down = (v, w, u) host arrays of the neighbouring slab (or None)."""
    def ptrs(arrs):
        if arrs is None:
            return None
        return (ctypes.c_void_p * 3)(*[a.ctypes.data for a in arrs])
    i2 = lambda x: (ctypes.c_int * 2)(*x)  # noqa: E731
    lib().fk_emu_set_mirror(ptrs(up), ptrs(down), i2(row0), i2(row1), i2(dst_row0))


def mirror_was_fused():
    return bool(lib().fk_emu_mirror_was_fused())


def plan_resident(H, W, batch=1, tiles=(0, 0), threads=0, nc=0, edge_tile=(0, 0), maps_global=0):
    """The resident kernel's geometry for a problem: dict or None."""
    out = (ctypes.c_int * 13)()
    force = (ctypes.c_int * 7)(int(tiles[0]), int(tiles[1]), threads, nc, int(edge_tile[0]), int(edge_tile[1]), maps_global)
    if not lib().fk_emu_plan_resident(H, W, batch, out, force):
        return None
    return dict(zip(("ntr", "ntc", "th_max", "tw_max", "threads", "smem_bytes", "nc", "xchg_bytes", "edge_rows",
                     "edge_colgroups", "single_phase", "maps_in_l2", "two_pass"), list(out)))


def plan_rows(H, W, batch=1, T=2, regs=168, top_off=0, bot_off=0):
    """Row-chunk boundaries of the streaming kernel's plan for a whole tissue: list of first rows, ending with H (or None)."""
    out = (ctypes.c_int * 4096)()
    if not lib().fk_emu_plan_rows(H, W, batch, T, regs, top_off, bot_off, out, 4096):
        return None
    return list(out[1:out[0] + 2])


def _pack_stims(stimuli):
    keep, arr = [], (_Stim * max(1, len(stimuli)))()
    for i, s in enumerate(stimuli):
        f = np.ascontiguousarray(s.field, dtype=np.float32)
        keep.append(f)
        arr[i] = _Stim(f.ctypes.data, *[float(np.asarray(x).reshape(-1)[0]) for x in s.protocol])
    return arr, keep


def dopri5(state, ts, params, D, stimuli, dx, rtol=1.4e-8, atol=1.4e-8, mxstep=float("inf"), exact=True):
    """The product's Dormand-Prince driver (fk_ode.h) + element bodies (fk_aux.h) on the CPU.  -> (v, w, u) stacked, stats"""
    v, w, u = [np.ascontiguousarray(x, dtype=np.float32) for x in state]
    H, W = u.shape
    ts = np.ascontiguousarray(ts, dtype=np.float32)
    D = np.ascontiguousarray(D, dtype=np.float32)
    par = np.array([float(np.asarray(x).reshape(-1)[0]) for x in params], dtype=np.float32)
    arr, keep = _pack_stims(stimuli)
    outs = [np.full((len(ts), H, W), np.nan, np.float32) for _ in range(3)]
    stats = (ctypes.c_longlong * 3)()
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib().fk_emu_dopri5(p(v), p(w), p(u), p(outs[0]), p(outs[1]), p(outs[2]), p(D), H, W, p(par), arr, len(stimuli),
                             p(ts), len(ts), ctypes.c_float(dx), ctypes.c_float(rtol), ctypes.c_float(atol),
                             ctypes.c_double(mxstep), int(exact), stats)
    if rc != 0:
        raise RuntimeError("fk_emu_dopri5 rc=%d" % rc)
    return tuple(outs), dict(attempts=stats[0], accepted=stats[1], rhs_evals=stats[2])


def resize(a, size):
    a = np.ascontiguousarray(a, dtype=np.float32)
    H, W = a.shape[-2:]
    planes = int(np.prod(a.shape[:-2], dtype=np.int64)) if a.ndim > 2 else 1
    out = np.empty(a.shape[:-2] + tuple(size), np.float32)
    lib().fk_emu_resize(a.ctypes.data_as(ctypes.c_void_p), planes, H, W, out.ctypes.data_as(ctypes.c_void_p), int(size[0]), int(size[1]))
    return out


def electrogram(x, point):
    x = np.ascontiguousarray(x, dtype=np.float32)
    H, W = x.shape[-2:]
    frames = int(np.prod(x.shape[:-2], dtype=np.int64)) if x.ndim > 2 else 1
    out = np.empty(x.shape[:-2], np.float32)
    lib().fk_emu_electrogram(x.ctypes.data_as(ctypes.c_void_p), frames, H, W, ctypes.c_float(point[0]), ctypes.c_float(point[1]),
                             out.ctypes.data_as(ctypes.c_void_p))
    return out


def heun(state, t0, t1, params, D, stimuli, dt, dx, exact=True):
    """The fused Heun driver (fk_driver.h: drive_heun, one tile launch per step) on the CPU.  -> (v, w, u), launches"""
    v, w, u = [np.ascontiguousarray(x, dtype=np.float32) for x in state]
    batch = u.shape[0] if u.ndim == 3 else 1
    H, W = u.shape[-2:]
    D = np.ascontiguousarray(D, dtype=np.float32)
    par = np.array([float(np.asarray(x).reshape(-1)[0]) for x in params], dtype=np.float32)
    arr, keep = _pack_stims(stimuli)
    vo, wo, uo = np.full_like(v, np.nan), np.full_like(w, np.nan), np.full_like(u, np.nan)
    n = ctypes.c_int(0)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib().fk_emu_heun(p(v), p(w), p(u), p(vo), p(wo), p(uo), p(D), H, W, batch, p(par), arr, len(stimuli),
                           ctypes.c_double(t0), ctypes.c_double(t1), ctypes.c_float(dt), ctypes.c_float(dx), int(exact),
                           ctypes.byref(n))
    if rc != 0:
        raise RuntimeError("fk_emu_heun rc=%d" % rc)
    return (vo, wo, uo), n.value


def heun_fast(state, t0, t1, params, D, stimuli, dt, dx, fold=True, force_stream=False):
    """The fast-numerics Heun driver (fk_driver.h: drive_heun_fast) on the CPU.  -> (v, w, u), dict of launch counts"""
    v, w, u = [np.ascontiguousarray(x, dtype=np.float32) for x in state]
    batch = u.shape[0] if u.ndim == 3 else 1
    H, W = u.shape[-2:]
    D = np.ascontiguousarray(D, dtype=np.float32)
    par = np.array([float(np.asarray(x).reshape(-1)[0]) for x in params], dtype=np.float32)
    arr, keep = _pack_stims(stimuli)
    vo, wo, uo = np.full_like(v, np.nan), np.full_like(w, np.nan), np.full_like(u, np.nan)
    info = (ctypes.c_int * 4)()
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib().fk_emu_heun_fast(p(v), p(w), p(u), p(vo), p(wo), p(uo), p(D), H, W, batch, p(par), arr, len(stimuli),
                                ctypes.c_double(t0), ctypes.c_double(t1), ctypes.c_float(dt), ctypes.c_float(dx), int(fold),
                                int(force_stream), info)
    if rc != 0:
        raise RuntimeError("fk_emu_heun_fast rc=%d" % rc)
    return (vo, wo, uo), dict(zip(("tile", "stream", "wide", "combine"), list(info)))
