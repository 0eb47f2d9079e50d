"""Loader and ctypes signatures of libfk.so (C ABI in include/fk.h).

The library is built in-tree (``cardiax_b200/csrc/libfk.so``) by ``build()`` with
``nvcc -gencode arch=compute_100a,code=sm_100a``.  There is no CPU fallback: every solver entry
point raises if the library or a CUDA device is missing.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.environ.get("FK_SO") or os.path.join(CSRC, "libfk.so")   # FK_SO: development A/B of two builds
_SOURCES = ["fk_api.cu", "fk_stream_tu.cu", "fk_resident.cu", "fk_aux.cu", "fk_core.h", "fk_tile.h", "fk_stream.h", "fk_stream.cuh",
            "fk_driver.h", "fk_wide.h", "fk_resident.h", "fk_resident.cuh", "fk_aux.h", "fk_aux.cuh", "fk_ode.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
STREAM_DEPTHS = (1, 2, 3, 4)   # one translation unit per (temporal-blocking depth, numerics), compiled in parallel

_lib = None


class FkParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_float) for n in (
        "tau_v_plus", "tau_v1_minus", "tau_v2_minus", "tau_w_plus", "tau_w_minus", "tau_d", "tau_0", "tau_r", "tau_si",
        "k", "V_csi", "V_c", "V_v", "Cm")]


class FkStimulus(ctypes.Structure):
    """include/fk.h: protocol values as doubles + which of them the caller held as integers (bits 1, 2, 4)."""
    _fields_ = [("field", ctypes.c_void_p), ("start", ctypes.c_double), ("duration", ctypes.c_double),
                ("period", ctypes.c_double), ("int_mask", ctypes.c_int), ("reserved", ctypes.c_int)]


class FkOptions(ctypes.Structure):
    _fields_ = [("exact", ctypes.c_int), ("steps_per_launch", ctypes.c_int), ("kernel", ctypes.c_int),
                ("phys_top", ctypes.c_int), ("phys_bottom", ctypes.c_int), ("cta_threads", ctypes.c_int),
                ("rows_per_cta", ctypes.c_int), ("uniform_diffusivity", ctypes.c_int), ("safe_division", ctypes.c_int),
                ("tiles_r", ctypes.c_int), ("tiles_c", ctypes.c_int), ("cells_per_thread", ctypes.c_int),
                ("edge_rows", ctypes.c_int), ("edge_colgroups", ctypes.c_int), ("maps_global", ctypes.c_int),
                ("counter_is_int", ctypes.c_int)]


class FkPeerMirror(ctypes.Structure):
    """include/fk.h: rows of a slab launch that are also stored into the neighbouring GPUs' (peer-mapped) arrays."""
    _fields_ = [("v", ctypes.c_void_p * 2), ("w", ctypes.c_void_p * 2), ("u", ctypes.c_void_p * 2),
                ("row0", ctypes.c_int * 2), ("row1", ctypes.c_int * 2), ("dst_row0", ctypes.c_int * 2)]


def _source_digest():
    """SHA-256 over the sources the library is built from (and the flags): what `libfk.so.sha` records at build time."""
    import hashlib
    h = hashlib.sha256()
    deps = [os.path.join(CSRC, s) for s in _SOURCES] + [os.path.join(_HERE, "..", "include", "fk.h")]
    for d in deps:
        h.update(os.path.basename(d).encode())
        if os.path.exists(d):
            with open(d, "rb") as f:
                h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(os.environ.get("FK_EXTRA_NVCC", "").encode())
    h.update(os.environ.get("FK_DEPTHS", "").encode())
    return h.hexdigest()


def needs_build():
    """Is the library missing or older than its sources?  Decided on the sources' CONTENT when the build left its digest
    next to the library (a checkout or a copy of the tree changes modification times, not what would be compiled);
    on modification times otherwise."""
    if not os.path.exists(SO_PATH):
        return True
    if os.environ.get("FK_SO"):   # a development A/B library, built with its own flags: never rebuilt behind the user's back
        return False
    sha = SO_PATH + ".sha"
    if os.path.exists(sha):
        try:
            with open(sha) as f:
                return f.read().strip() != _source_digest()
        except OSError:
            pass
    t = os.path.getmtime(SO_PATH)
    deps = [os.path.join(CSRC, s) for s in _SOURCES] + [os.path.join(_HERE, "..", "include", "fk.h")]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """nvcc-compile csrc/*.cu into csrc/libfk.so (cross-compiles without a GPU): fk_api.cu and one object of
    fk_stream_tu.cu per (temporal-blocking depth, numerics), in parallel, then one link."""
    if not force and not needs_build():
        return SO_PATH
    # One builder at a time: under torchrun every rank may find the library stale at once (a source file touched after the
    # last build), and eight concurrent links into the same file left a truncated libfk.so behind.  The ranks queue on a
    # lock file; whoever gets it second finds the library fresh and returns.  The link goes to a temporary name and is
    # renamed into place, so a reader never sees a half-written file.
    import fcntl
    os.makedirs(os.path.dirname(SO_PATH), exist_ok=True)
    with open(SO_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():
                return SO_PATH
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force, verbose):
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC") or ("/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else "nvcc")
    env = dict(os.environ)
    env.pop("CC", None)   # the image's CC/CXX point at a gcc without its support files
    env.pop("CXX", None)
    objdir = os.environ.get("FK_OBJDIR") or os.path.join(CSRC, "build")   # FK_OBJDIR + FK_SO: a second build next to the first
    os.makedirs(objdir, exist_ok=True)
    depths = STREAM_DEPTHS
    if os.environ.get("FK_DEPTHS"):   # development: e.g. FK_DEPTHS=2 links only the T = 2 kernels (others: "unsupported")
        depths = tuple(int(x) for x in os.environ["FK_DEPTHS"].split(","))
    mask = sum(1 << t for t in depths)
    jobs = [(os.path.join(objdir, "fk_api.o"), ["-DFK_DEPTH_MASK=%d" % mask, "-c", os.path.join(CSRC, "fk_api.cu")])]
    jobs.append((os.path.join(objdir, "fk_resident.o"), ["-c", os.path.join(CSRC, "fk_resident.cu")]))
    jobs.append((os.path.join(objdir, "fk_aux.o"), ["-c", os.path.join(CSRC, "fk_aux.cu")]))
    for t in reversed(depths):   # the deepest (slowest to compile) first
        for e in (1, 0):
            jobs.append((os.path.join(objdir, "fk_stream_T%d_E%d.o" % (t, e)),
                         ["-DFK_TU_T=%d" % t, "-DFK_TU_EXACT=%d" % e, "-c", os.path.join(CSRC, "fk_stream_tu.cu")]))
    ptxas = ["-Xptxas", "-v"] if verbose else []
    ptxas += os.environ.get("FK_EXTRA_NVCC", "").split()   # development: e.g. FK_EXTRA_NVCC="-DFK_MAP_LATE=1" for an A/B build
    aux_deps = ["fk_aux.cu", "fk_aux.cuh", "fk_aux.h", "fk_ode.h", "fk_core.h"]
    res_deps = [d for d in _SOURCES if d not in ("fk_aux.cu", "fk_aux.cuh", "fk_aux.h", "fk_ode.h")]
    stream_deps = ["fk_stream_tu.cu", "fk_stream.cuh", "fk_stream.h", "fk_tile.h", "fk_core.h"]

    def fresh(job):   # an object newer than everything its translation unit includes is kept (FK_DEPTHS changes: force)
        obj, args = job
        if force or not os.path.exists(obj) or "fk_api.cu" in args[-1]:
            return False
        deps = stream_deps if "fk_stream_tu.cu" in args[-1] else aux_deps if "fk_aux.cu" in args[-1] else res_deps
        return all(os.path.getmtime(obj) >= os.path.getmtime(os.path.join(CSRC, d)) for d in deps)

    def run(job):
        obj, args = job
        if fresh(job):
            return ""
        out = subprocess.run([nvcc] + NVCC_FLAGS + ptxas + args + ["-o", obj], capture_output=True, text=True, env=env)
        if out.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + out.stdout + out.stderr)
        return out.stderr

    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as pool:
        logs = list(pool.map(run, jobs))
    tmp = SO_PATH + ".tmp.%d" % os.getpid()
    out = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", tmp] + [j[0] for j in jobs],
                         capture_output=True, text=True, env=env)
    if out.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("nvcc link failed:\n" + out.stdout + out.stderr)
    digest = _source_digest()
    os.replace(tmp, SO_PATH)
    with open(SO_PATH + ".sha", "w") as f:
        f.write(digest + "\n")
    if verbose:
        print("\n".join(logs))
    return SO_PATH


def lib():
    """The loaded library; raises RuntimeError if it is not built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError("cardiax_b200: %s is missing -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a).  There is no CPU fallback." % SO_PATH)
    L = ctypes.CDLL(SO_PATH)
    vp, ci, cf, cd, ll, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_longlong, ctypes.c_size_t
    L.fk_abi_version.restype = ci
    L.fk_last_error.restype = ctypes.c_char_p
    L.fk_default_options.argtypes = [ctypes.POINTER(FkOptions)]
    L.fk_default_options.restype = None
    L.fk_workspace_bytes.argtypes = [ci, ci, ci, ci, ci]
    L.fk_workspace_bytes.restype = sz
    L.fk_forward_euler.argtypes = [vp] * 6 + [vp, ci, ci, ci, ci, ctypes.POINTER(FkParams), ctypes.POINTER(FkStimulus), ci,
                                              cd, cd, cf, cf, ctypes.POINTER(FkOptions), vp, sz, vp]
    L.fk_forward_euler.restype = ci
    L.fk_forward_heun.argtypes = L.fk_forward_euler.argtypes
    L.fk_forward_heun.restype = ci
    L.fk_heun_workspace_bytes.argtypes = [ci, ci, ci, ci, ci]
    L.fk_heun_workspace_bytes.restype = sz
    L.fk_euler_rows.argtypes = [vp] * 6 + [vp, vp, vp, ci, ci, ctypes.POINTER(FkParams), ctypes.POINTER(FkStimulus), ci, cd,
                                           ci, cf, cf, ctypes.POINTER(FkOptions), ci, ci, vp, sz, vp]
    L.fk_euler_rows.restype = ci
    L.fk_euler_rows_peer.argtypes = L.fk_euler_rows.argtypes + [ctypes.POINTER(FkPeerMirror), ctypes.POINTER(ci)]
    L.fk_euler_rows_peer.restype = ci
    L.fk_peer_alloc.argtypes = [sz, ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_ubyte * 64)]
    L.fk_peer_open.argtypes = [ctypes.POINTER(ctypes.c_ubyte * 64), ctypes.POINTER(vp)]
    L.fk_peer_close.argtypes = [vp]
    L.fk_peer_free.argtypes = [vp]
    L.fk_peer_signal.argtypes = [vp, ctypes.c_uint, vp]
    L.fk_peer_wait.argtypes = [vp, ctypes.c_uint, vp]
    L.fk_peer_copy.argtypes = [vp, vp, sz, vp]
    for f in (L.fk_peer_alloc, L.fk_peer_open, L.fk_peer_close, L.fk_peer_free, L.fk_peer_signal, L.fk_peer_wait, L.fk_peer_copy):
        f.restype = ci
    L.fk_check_exact_division.argtypes = [ctypes.POINTER(FkParams), cf, ctypes.POINTER(ll), vp]
    L.fk_check_exact_division.restype = ci
    L.fk_rhs.argtypes = [vp] * 6 + [vp, ci, ci, ci, ci, ctypes.POINTER(FkParams), ctypes.POINTER(FkStimulus), ci, cd, cf,
                                    ctypes.POINTER(FkOptions), vp, sz, vp]
    L.fk_rhs.restype = ci
    L.fk_gradient.argtypes = [vp, vp, ll, ll, ll, vp]
    L.fk_gradient.restype = ci
    L.fk_stimulate.argtypes = [cd, ci, vp, vp, ci, ci, ctypes.POINTER(FkStimulus), ci, vp, sz, vp]
    L.fk_stimulate.restype = ci
    L.fk_diffusivity_gradients.argtypes = [vp, vp, vp, ci, ci, ci, cf, ci, ci, vp]
    L.fk_diffusivity_gradients.restype = ci
    L.fk_dopri5_workspace_bytes.argtypes = [ci, ci, ci, ci, ci]
    L.fk_dopri5_workspace_bytes.restype = sz
    L.fk_odeint_dopri5.argtypes = [vp] * 6 + [vp, ci, ci, ci, ci, ctypes.POINTER(FkParams), ctypes.POINTER(FkStimulus), ci,
                                              ctypes.POINTER(cf), ci, cf, cf, cf, cd, ctypes.POINTER(FkOptions), vp, sz, vp,
                                              ctypes.POINTER(ll * 3)]
    L.fk_odeint_dopri5.restype = ci
    L.fk_resize_workspace_bytes.argtypes = [ci, ci, ci, ci, ci]
    L.fk_resize_workspace_bytes.restype = sz
    L.fk_resize_bilinear.argtypes = [ctypes.POINTER(vp), ci, ci, ci, vp, ci, ci, vp, sz, ci, vp]
    L.fk_resize_bilinear.restype = ci
    L.fk_electrogram.argtypes = [vp, ci, ci, ci, cf, cf, vp, vp]
    L.fk_electrogram.restype = ci
    L.fk_launch_count.restype = ll
    L.fk_last_plan.argtypes = [ctypes.POINTER(ci * 8)]
    L.fk_last_plan.restype = None
    L.fk_last_kernel.restype = ctypes.c_char_p
    L.fk_resident_timing.argtypes = [ctypes.POINTER(ctypes.c_ulonglong * 8)]
    L.fk_resident_timing.restype = ci
    L.fk_profile_enable.argtypes = [ci]
    L.fk_profile_enable.restype = None
    L.fk_profile_collect.argtypes = [ctypes.POINTER(cd), ctypes.POINTER(ll), ctypes.POINTER(cd), ctypes.POINTER(ll),
                                      ctypes.POINTER(cd)]
    L.fk_profile_collect.restype = ci
    L.fk_profile_dropped.restype = ll
    if L.fk_abi_version() != 2:
        raise RuntimeError("libfk.so ABI version mismatch")
    _lib = L
    return L


def last_kernel():
    return lib().fk_last_kernel().decode()


def last_plan():
    out = (ctypes.c_int * 8)()
    lib().fk_last_plan(ctypes.byref(out))
    if last_kernel() in ("fk_resident_kernel", "fk_cluster_kernel"):
        d = dict(zip(("steps", "cta_threads", "tile_cols", "tile_w", "tile_h", "tile_rows", "cells_per_thread", "smem_bytes"),
                     list(out)))
        d["maps_in_l2"], d["cells_per_thread"] = d["cells_per_thread"] >> 3, d["cells_per_thread"] & 7
        return d
    return dict(zip(("T", "cta_threads", "strips", "cols_per_strip", "rows_per_cta", "row_chunks", "ctas_per_sm",
                     "smem_bytes"), list(out)))


def check(rc):
    if rc != 0:
        msg = lib().fk_last_error().decode()
        if rc < 0:
            raise ValueError("libfk: %s (code %d)" % (msg, rc))
        raise RuntimeError("libfk: %s (cudaError %d)" % (msg, rc))
