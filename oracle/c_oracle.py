"""ctypes binding of the C restatement (oracle/fk_oracle.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfk_oracle.so")
_lib = None


def build(force=False):
    """Compile oracle/fk_oracle.c with gcc (no GPU, no reference files needed)."""
    srcs = [os.path.join(_HERE, f) for f in ("fk_oracle.c", "fk_oracle_impl.h")]
    def fresh():
        return os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs)
    if not force and fresh():
        return _SO
    import fcntl   # (one builder at a time: several test / bench processes may find the library stale at once)
    with open(_SO + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if force or not fresh():
                subprocess.check_call(["make", "-C", _HERE, "-B", "libfk_oracle.so"], stdout=subprocess.DEVNULL)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        for suf, ct in (("f32", ctypes.c_float), ("f64", ctypes.c_double)):
            fn = getattr(_lib, "fk_oracle_forward_euler_" + suf)
            p = ctypes.POINTER(ct)
            fn.argtypes = [p, p, p, p, ctypes.c_long, ctypes.c_long, p, ctypes.POINTER(p), p, ctypes.c_int,
                           ctypes.c_double, ctypes.c_long, ct, ct, ctypes.c_int]
            fn.restype = ctypes.c_int
            fa = getattr(_lib, "fk_oracle_stim_active_" + suf)
            fa.argtypes = [ct, ct, ct, ct]
            fa.restype = ctypes.c_int
        _lib.fk_oracle_set_threads.argtypes = [ctypes.c_int]
        _lib.fk_oracle_set_threads.restype = ctypes.c_int
    return _lib


def set_threads(n):
    return lib().fk_oracle_set_threads(int(n))


def forward_euler(state, t, t_end, params, diffusivity, stimuli, dt, dx, dtype=np.float32, tanh="xla"):
    """Same contract as oracle.fk_oracle.forward_euler (cardiax/solve.py:92-100); returns a new State."""
    from .fk_oracle import State
    L = lib()
    f32 = np.dtype(dtype) == np.float32
    ct = ctypes.c_float if f32 else ctypes.c_double
    p = ctypes.POINTER(ct)
    fn = L.fk_oracle_forward_euler_f32 if f32 else L.fk_oracle_forward_euler_f64
    v, w, u = [np.array(x, dtype=dtype, order="C", copy=True) for x in state]
    D = np.ascontiguousarray(diffusivity, dtype=dtype)
    H, W = u.shape
    par = np.array([float(np.asarray(x).reshape(-1)[0]) for x in params], dtype=dtype)
    fields = [np.ascontiguousarray(s.field, dtype=dtype) for s in stimuli]
    proto = np.array([[float(np.asarray(x).reshape(-1)[0]) for x in s.protocol] for s in stimuli],
                     dtype=dtype).reshape(-1)
    arr = (p * max(1, len(fields)))(*[f.ctypes.data_as(p) for f in fields])
    nsteps = 0
    i = float(t)
    while i < float(t_end):
        nsteps += 1
        i += 1.0
    rc = fn(v.ctypes.data_as(p), w.ctypes.data_as(p), u.ctypes.data_as(p), D.ctypes.data_as(p), H, W,
            par.ctypes.data_as(p), arr, proto.ctypes.data_as(p) if len(fields) else None, len(fields),
            float(t), nsteps, ct(dt), ct(dx), 0 if tanh == "xla" else 1)
    if rc != 0:
        raise MemoryError("fk_oracle_forward_euler failed")
    return State(v, w, u)
