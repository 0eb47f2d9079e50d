"""Run the UNMODIFIED reference sources (``/root/reference/cardiax/{solve,stimulus,params,convert}.py``) on NumPy.

TEST INFRASTRUCTURE ONLY.  The reference is pure Python on top of a 2021 ``jax`` that cannot be installed here
(SURVEY 8c).  This module is a NumPy-backed stand-in for the handful of ``jax`` entry points those four files touch, so
that the reference's own source text -- loaded by path, not copied, not edited -- executes and its results can be
compared with ``oracle/`` bit for bit (``tests/test_reference_pin.py``) and frozen as golden vectors
(``tests/golden/make_reference_golden.py``).

What the stand-in has to get right is jax's *semantics* with x64 disabled, because they decide where values round:

* there are no 64-bit types: arrays are bool / int32 / float32;
* type promotion follows jax's lattice  bool < int < float  at 32 bits, Python scalars are weakly typed (they take
  the array's kind or lift it, never its width): ``(1 - q) * 19.6`` with a bool ``q`` is float32, ``p / 50`` is float32;
* ``jax.jit`` hands every non-static Python scalar argument (and every leaf of a NamedTuple / list argument) to the
  function as a 32-bit scalar, so scalar-scalar arithmetic inside a jitted function is 32-bit as well;
* ``lax.fori_loop(lower, upper, body, init)`` runs ``i = lower; while i < upper: val = body(i, val); i = i + 1`` with a
  counter of the bounds' dtype -- float32 through ``solve.forward`` (it passes ``float(checkpoint)``), int32 through
  ``deepx.generate.sequence``;
* ``jnp.mod`` is ``lax.rem`` plus the sign-of-the-divisor correction; ``jnp.where`` tests ``cond != 0``.

What it cannot reproduce is XLA's code generation below the op level: whether ``a * b + c`` is contracted and which
``tanh`` it emits.  ``tanh`` is therefore a switch: ``"numpy"`` (``np.tanh``) or ``"xla"`` (the published fp32 rational of
``xla/service/llvm_ir/math_ops.cc``: EmitFastTanh, jaxlib 0.1.64).  Every other operation is one IEEE fp32 operation per
jax op, un-contracted.

Usage::

    ref = load_reference()            # ref.solve, ref.stimulus, ref.params, ref.convert: the reference's modules
    with ref.tanh("xla"): ref.solve._forward_euler(...)
"""
import contextlib
import functools
import hashlib
import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("FK_REFERENCE_ROOT", "/root/reference")
REFERENCE_FILES = ("cardiax/solve.py", "cardiax/stimulus.py", "cardiax/params.py", "cardiax/convert.py")

# sha256 of the reference files this shim was validated against (epignatelli/cardiax as mounted at /root/reference)
REFERENCE_SHA256 = {
    "cardiax/solve.py": "dce12fd163e797d421f90c432d4d4335121a96b1f4b61c6543b854fca3fedfe0",
    "cardiax/stimulus.py": "4fb303230c1f6347c9c8e90762eace254f7a7c0bd0344353bd76eab3ffc05ba0",
    "cardiax/params.py": "62ecf6c178678b57bd93b943906f049b142967420321e80d78078f2072241e91",
    "cardiax/convert.py": "8821de98d4e6aec721b1dbd62c88213289e105f7ffced27aea4cdc87894440dd",
}

_B, _I, _F = 0, 1, 2
_DTYPES = {_B: np.bool_, _I: np.int32, _F: np.float32}
_state = {"tanh": "numpy"}


# ----------------------------------------------------------------------------------------------- array type
def _kind(x):
    if isinstance(x, (bool, np.bool_)):
        return _B
    if isinstance(x, (int, np.integer)):
        return _I
    if isinstance(x, (float, np.floating)):
        return _F
    k = np.asarray(x).dtype.kind
    return _B if k == "b" else _I if k in "iu" else _F


def _raw(x, kind):
    """Plain ndarray of the 32-bit dtype of ``kind`` (int -> float conversions round to nearest, like XLA)."""
    return np.asarray(x).view(np.ndarray).astype(_DTYPES[kind], copy=False) if isinstance(x, np.ndarray) \
        else np.asarray(x, dtype=_DTYPES[kind])


def _wrap(x):
    x = np.asarray(x)
    k = _kind(x)
    if x.dtype != _DTYPES[k]:
        x = x.astype(_DTYPES[k])
    return x.view(JArray)


_COMPARE = {np.greater_equal, np.greater, np.less, np.less_equal, np.equal, np.not_equal}
_FLOAT_RESULT = {np.true_divide, np.tanh, np.exp, np.sqrt}
_SAME_KIND = {np.add, np.subtract, np.multiply, np.negative, np.absolute, np.minimum, np.maximum, np.bitwise_and,
              np.bitwise_or, np.logical_and, np.logical_or, np.logical_not, np.fmod, np.power, np.positive}


class JArray(np.ndarray):
    """ndarray with jax's x64-disabled promotion: operands are brought to ONE 32-bit dtype, then one NumPy op runs."""
    __array_priority__ = 100

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        if method != "__call__" or out is not None or kwargs:
            raise NotImplementedError("ref_shim: %s.%s(out=%s, %s) is not modelled" % (ufunc, method, out, kwargs))
        kind = max(_kind(x) for x in inputs)
        if ufunc in _FLOAT_RESULT:
            kind = _F
        elif ufunc not in _COMPARE and ufunc not in _SAME_KIND:
            raise NotImplementedError("ref_shim: ufunc %s is not modelled" % ufunc)
        args = [_raw(x, kind) for x in inputs]
        if ufunc is np.tanh:
            return _wrap(_tanh(args[0]))
        with np.errstate(all="ignore"):
            return _wrap(ufunc(*args))

    # Python's in-place operators would ask NumPy for out=self; jax arrays are immutable, `a &= b` rebinds
    def __iand__(self, other):
        return self & other

    def __ior__(self, other):
        return self | other

    def __iadd__(self, other):
        return self + other

    def __isub__(self, other):
        return self - other

    def __imul__(self, other):
        return self * other

    def __itruediv__(self, other):
        return self / other

    def __bool__(self):
        return bool(np.asarray(self).reshape(-1)[0]) if self.size == 1 else np.ndarray.__bool__(self)

    def __index__(self):
        if self.size == 1 and self.dtype.kind in "iu":
            return int(np.asarray(self).reshape(-1)[0])
        raise TypeError("only integer scalar arrays can be converted to an index")

    def astype(self, dtype, *a, **k):     # `checkpoints.astype(float)`: float is float32 with x64 disabled
        return _wrap(np.asarray(self).astype(_DTYPES[_kind(np.zeros((), dtype))]))


# XLA's fp32 tanh, third-party (jaxlib 0.1.64 pinned by the reference's install_jax.sh:2): EmitFastTanh clamps to
# [-9, 9], evaluates an odd degree-13 numerator and an even degree-6 denominator in x^2 by Horner's rule, and returns
# x itself below 0.0004.  One fp32 multiply and one fp32 add per Horner step, no contraction.
_NUM = (-2.76076847742355e-16, 2.00018790482477e-13, -8.60467152213735e-11, 5.12229709037114e-08,
        1.48572235717979e-05, 6.37261928875436e-04, 4.89352455891786e-03)
_DEN = (1.19825839466702e-06, 1.18534705686654e-04, 2.26843463243900e-03, 4.89352518554385e-03)


def _tanh(x):
    if _state["tanh"] == "numpy":
        return np.tanh(x)
    f = np.float32
    c = np.clip(x, f(-9.0), f(9.0))
    s = c * c
    n = np.full_like(s, f(_NUM[0]))
    for a in _NUM[1:]:
        n = n * s + f(a)
    d = np.full_like(s, f(_DEN[0]))
    for a in _DEN[1:]:
        d = d * s + f(a)
    return np.where(np.abs(x) < f(0.0004), x, (c * n) / d)


# ----------------------------------------------------------------------------------------------- jax.numpy
def _promote(*xs):
    kind = max(_kind(x) for x in xs)
    return [_raw(x, kind) for x in xs]


def _mod(a, b):
    """jnp.mod = lax.rem (C fmod / truncated integer remainder) + sign-of-the-divisor correction."""
    a, b = _promote(a, b)
    with np.errstate(all="ignore"):
        r = np.fmod(a, b)
        fix = (r != 0) & ((r < 0) != (b < 0))
        return _wrap(np.where(fix, r + b, r))


def _where(c, a, b):
    a, b = _promote(a, b)
    return _wrap(np.where(np.asarray(c).view(np.ndarray) != 0, a, b))


def _default_dtype(dtype, fallback=np.float32):
    return _DTYPES[_kind(np.zeros((), dtype))] if dtype is not None else fallback


def _arange(*args, dtype=None):
    vals = [a.item() if isinstance(a, np.ndarray) else a for a in args]
    kind = max(_kind(v) for v in vals)
    return _wrap(np.arange(*vals).astype(_default_dtype(dtype, _DTYPES[kind])))


def _array(x, dtype=None):
    a = np.asarray(x)
    return _wrap(a if dtype is None else a.astype(_default_dtype(dtype)))


def _make_jnp():
    m = types.ModuleType("jax.numpy")
    m.ndarray = JArray
    m.float32, m.int32, m.bool_ = np.float32, np.int32, np.bool_
    m.pad = lambda a, w, mode="constant": _wrap(np.pad(np.asarray(a).view(np.ndarray), w, mode=mode))
    m.concatenate = lambda xs, axis=0: _wrap(np.concatenate(_promote(*xs), axis))
    m.where = _where
    m.mod = _mod
    m.zeros = lambda shape, dtype=None: _wrap(np.zeros(shape, _default_dtype(dtype)))
    m.ones = lambda shape, dtype=None: _wrap(np.ones(shape, _default_dtype(dtype)))
    m.zeros_like = lambda a: _wrap(np.zeros_like(np.asarray(a).view(np.ndarray)))
    m.arange = _arange
    m.array = _array
    m.asarray = _array
    for name in ("add", "subtract", "multiply", "greater_equal", "greater", "less", "tanh", "abs", "minimum", "maximum"):
        m.__dict__[name] = (lambda uf: lambda *a: uf(*[x if isinstance(x, JArray) else _wrap(x) for x in a]))(
            getattr(np, name))
    m.sum = lambda a, *args, **kw: _wrap(np.sum(np.asarray(a).view(np.ndarray), *args, **kw))
    m.mean = lambda a, *args, **kw: _wrap(np.mean(np.asarray(a).view(np.ndarray), *args, **kw))
    m.nonzero = lambda a: tuple(_wrap(i) for i in np.nonzero(np.asarray(a)))
    return m


# ----------------------------------------------------------------------------------------------- jax / jax.lax / jax.ops
def _is_namedtuple(x):
    return isinstance(x, tuple) and hasattr(x, "_fields")


def _tree_map(f, tree, *rest):
    if _is_namedtuple(tree):
        return type(tree)(*[_tree_map(f, t, *[r[i] for r in rest]) for i, t in enumerate(tree)])
    if isinstance(tree, (list, tuple)):
        return type(tree)(_tree_map(f, t, *[r[i] for r in rest]) for i, t in enumerate(tree))
    if isinstance(tree, dict):
        return {k: _tree_map(f, v, *[r[k] for r in rest]) for k, v in tree.items()}
    return f(tree, *rest)


def _trace_leaf(x):
    """What a jitted function sees for a leaf of its arguments: a 32-bit array (Python scalars included)."""
    if x is None or isinstance(x, (str, types.FunctionType)):
        return x
    return _wrap(x)


def _jit(fun=None, static_argnums=()):
    if fun is None:
        return functools.partial(_jit, static_argnums=static_argnums)
    static = (static_argnums,) if isinstance(static_argnums, int) else tuple(static_argnums)

    @functools.wraps(fun)
    def jitted(*args, **kwargs):
        args = [a if i in static else _tree_map(_trace_leaf, a) for i, a in enumerate(args)]
        return fun(*args, **{k: _tree_map(_trace_leaf, v) for k, v in kwargs.items()})

    jitted.__wrapped_reference__ = fun
    return jitted


def _fori_loop(lower, upper, body_fun, init_val):
    lower, upper = _promote(lower, upper)
    one = lower.dtype.type(1)
    i = _wrap(lower)
    val = init_val
    while bool(np.asarray(i) < upper):
        val = body_fun(i, val)
        i = i + one
    return val


def _slice_in_dim(a, start_index, limit_index, stride=1, axis=0):
    idx = [slice(None)] * np.ndim(a)
    idx[axis] = slice(start_index, limit_index, stride)     # Python slice semantics, negative indices included
    return _wrap(np.asarray(a).view(np.ndarray)[tuple(idx)])


class _IndexHelper:
    def __getitem__(self, idx):
        return idx


def _index_update(x, idx, y):
    out = np.array(np.asarray(x).view(np.ndarray), copy=True)
    if not isinstance(idx, tuple):
        idx = (idx,)
    idx = tuple(slice(*[None if v is None else int(np.asarray(v).reshape(-1)[0]) for v in (s.start, s.stop, s.step)])
                if isinstance(s, slice) else s for s in idx)
    out[idx] = np.asarray(y, dtype=out.dtype) if np.ndim(y) else out.dtype.type(np.asarray(y).reshape(-1)[0])
    return _wrap(out)


def _unavailable(name):
    def f(*a, **k):
        raise NotImplementedError("ref_shim: %s is an un-vendored jax component and is not modelled" % name)
    return f


def _make_modules():
    jnp = _make_jnp()
    jax = types.ModuleType("jax")
    jax.__path__ = []
    jax.numpy = jnp
    jax.jit = _jit
    jax.tree_multimap = _tree_map
    jax.tree_map = _tree_map
    lax = types.ModuleType("jax.lax")
    lax.fori_loop = _fori_loop
    lax.slice_in_dim = _slice_in_dim
    jax.lax = lax
    ops = types.ModuleType("jax.ops")
    ops.index = _IndexHelper()
    ops.index_update = _index_update
    jax.ops = ops
    exp = types.ModuleType("jax.experimental")
    exp.__path__ = []
    ode = types.ModuleType("jax.experimental.ode")
    ode.odeint = _unavailable("jax.experimental.ode.odeint")
    exp.ode = ode
    jax.experimental = exp
    mpl = types.ModuleType("matplotlib")
    mpl.__path__ = []
    plt = types.ModuleType("matplotlib.pyplot")
    plt.show = lambda *a, **k: None
    mpl.pyplot = plt
    return {"jax": jax, "jax.numpy": jnp, "jax.lax": lax, "jax.ops": ops, "jax.experimental": exp,
            "jax.experimental.ode": ode, "matplotlib": mpl, "matplotlib.pyplot": plt}


# ----------------------------------------------------------------------------------------------- loader
def available(root=None):
    root = root or REFERENCE_ROOT
    return all(os.path.isfile(os.path.join(root, f)) for f in REFERENCE_FILES)


def file_hashes(root=None):
    root = root or REFERENCE_ROOT
    out = {}
    for f in REFERENCE_FILES:
        with open(os.path.join(root, f), "rb") as fh:
            out[f] = hashlib.sha256(fh.read()).hexdigest()
    return out


class Reference:
    """The reference's modules, executing on the stand-in."""

    def __init__(self, solve, stimulus, params, convert):
        self.solve, self.stimulus, self.params, self.convert = solve, stimulus, params, convert

    @staticmethod
    @contextlib.contextmanager
    def tanh(which):
        assert which in ("numpy", "xla")
        old, _state["tanh"] = _state["tanh"], which
        try:
            yield
        finally:
            _state["tanh"] = old


_loaded = {}


def load_reference(root=None):
    """Import the four reference files by path, unmodified, under a private package name; `jax` and `matplotlib`
    resolve to the stand-ins only while those files are being imported (sys.modules is restored afterwards)."""
    root = root or REFERENCE_ROOT
    if root in _loaded:
        return _loaded[root]
    if not available(root):
        raise FileNotFoundError("reference sources not found under %s" % root)
    fakes = _make_modules()
    pkg_name = "_fk_reference_cardiax"
    saved = {k: sys.modules.get(k) for k in list(fakes) + [pkg_name]}
    import warnings
    try:
        sys.modules.update(fakes)
        pkg = types.ModuleType(pkg_name)
        pkg.__path__ = []
        sys.modules[pkg_name] = pkg
        plot = types.ModuleType(pkg_name + ".plot")     # cardiax/plot.py is matplotlib drawing: out of the path
        sys.modules[pkg_name + ".plot"] = plot
        pkg.plot = plot
        mods = {}
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", DeprecationWarning)   # scipy.ndimage.interpolation (stimulus.py:7)
            for name in ("convert", "params", "stimulus", "solve"):
                spec = importlib.util.spec_from_file_location("%s.%s" % (pkg_name, name),
                                                              os.path.join(root, "cardiax", name + ".py"))
                mod = importlib.util.module_from_spec(spec)
                sys.modules[spec.name] = mod
                spec.loader.exec_module(mod)
                setattr(pkg, name, mod)
                mods[name] = mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        for k in [k for k in sys.modules if k.startswith(pkg_name)]:
            sys.modules.pop(k, None)
    ref = Reference(mods["solve"], mods["stimulus"], mods["params"], mods["convert"])
    _loaded[root] = ref
    return ref


def to_numpy(tree):
    """State / list of States of JArrays -> plain float32 ndarrays."""
    return _tree_map(lambda x: np.array(np.asarray(x).view(np.ndarray), copy=True), tree)
