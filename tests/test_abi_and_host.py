"""No-GPU checks: libfk.so loads and exports every symbol include/fk.h declares; host-side mirrors of the reference's
setup code (params, stimulus builders, convert) agree with the oracle / the reference's formulas."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from cardiax_b200 import _lib
    _lib.build()
    hdr = open(os.path.join(ROOT, "include", "fk.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(fk_[a-z0-9_]+)\s*\(", hdr))
    assert {"fk_forward_euler", "fk_rhs", "fk_gradient", "fk_stimulate", "fk_workspace_bytes",
            "fk_diffusivity_gradients", "fk_forward_euler_slab" if False else "fk_abi_version"} <= names
    L = ctypes.CDLL(_lib.SO_PATH)
    for n in sorted(names):
        assert hasattr(L, n), n
    assert _lib.lib().fk_abi_version() == 2
    # ABI v2: the protocol travels as doubles + an integer-typing mask ("f64 or i64 + flag")
    assert ctypes.sizeof(_lib.FkParams) == 56 and ctypes.sizeof(_lib.FkStimulus) == 40 and ctypes.sizeof(_lib.FkOptions) == 64
    assert ctypes.sizeof(_lib.FkPeerMirror) == 72
    assert {"fk_euler_rows_peer", "fk_peer_alloc", "fk_peer_open", "fk_peer_signal", "fk_peer_wait", "fk_peer_copy"} <= names
    # argument errors are reported, not crashed on (no GPU needed: validation comes first)
    P = _lib.FkParams(*([1.0] * 14))
    rc = _lib.lib().fk_forward_euler(None, None, None, None, None, None, None, 0, 2, 2, 1, ctypes.byref(P), None, 0, 0.0, 1.0,
                                     0.01, 0.01, None, None, 0, None)
    assert rc < 0 and b"3 x 3" in _lib.lib().fk_last_error()
    assert _lib.lib().fk_workspace_bytes(4096, 4096, 1, 0, 0) >= 5 * 4096 * 4096 * 4


def test_new_entry_points_validate_their_arguments_before_touching_the_device():
    """fk_odeint_dopri5 / fk_resize_bilinear / fk_electrogram (the SURVEY 8f rows): argument errors come back as negative
    codes with a message, like the rest of the ABI; the workspace queries are pure host arithmetic."""
    from cardiax_b200 import _lib
    L = _lib.lib()
    P = _lib.FkParams(*([1.0] * 14))
    ts = (ctypes.c_float * 2)(0.0, 1.0)
    rc = L.fk_odeint_dopri5(None, None, None, None, None, None, None, 0, 16, 16, 1, ctypes.byref(P), None, 0, ts, 2, 0.01, 1e-5,
                            1e-5, float("inf"), None, None, 0, None, None)
    assert rc < 0 and b"NULL" in L.fk_last_error()
    rc = L.fk_odeint_dopri5(None, None, None, None, None, None, None, 0, 2, 2, 1, ctypes.byref(P), None, 0, ts, 2, 0.01, 1e-5,
                            1e-5, float("inf"), None, None, 0, None, None)
    assert rc < 0 and b"3 x 3" in L.fk_last_error()
    base = L.fk_workspace_bytes(64, 96, 1, 2, 0)
    assert L.fk_dopri5_workspace_bytes(64, 96, 1, 2, 0) >= base + 60 * 64 * 96 * 4      # 20 States of scratch
    assert L.fk_heun_workspace_bytes(64, 96, 1, 2, 0) >= base + 9 * 64 * 96 * 4
    assert L.fk_resize_bilinear(None, 3, 8, 8, None, 4, 4, None, 0, 0, None) < 0 and b"NULL" in L.fk_last_error()
    one = (ctypes.c_void_p * 1)(8)
    assert L.fk_resize_bilinear(one, 1, 0, 8, ctypes.c_void_p(8), 4, 4, ctypes.c_void_p(8), 1 << 20, 0, None) < 0
    assert b"shape" in L.fk_last_error()
    assert L.fk_resize_bilinear(one, 1, 8, 8, ctypes.c_void_p(8), 4, 4, ctypes.c_void_p(8), 16, 0, None) == -4   # workspace too small
    need = L.fk_resize_workspace_bytes(1200, 1200, 256, 256, 3)
    # two index tables + tap-major weights: K = floor(2 * 1200 / 256) + 2 = 11 taps per output index and axis
    assert 2 * (256 * 4 + 256 * 11 * 4) <= need <= 2 * (256 * 4 + 256 * 11 * 4) + 5 * 256
    assert L.fk_resize_workspace_bytes(0, 8, 4, 4, 1) == 0
    assert L.fk_electrogram(None, 1, 8, 8, 0.0, 0.0, None, None) < 0
    assert L.fk_electrogram(ctypes.c_void_p(8), 0, 8, 8, 0.0, 0.0, ctypes.c_void_p(8), None) < 0


def test_solver_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cardiax_b200 import solve
    with pytest.raises(RuntimeError):
        solve.init((8, 8))
    with pytest.raises(RuntimeError):
        solve._forward_euler(O.init((8, 8)), 0, 1, O.PARAMSETS["3"], np.ones((8, 8), np.float32), [], 0.01, 0.01)


def test_product_never_imports_the_oracle():
    for base in ("cardiax_b200", "cardiax", "deepx"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".h", ".cuh")):
                    src = open(os.path.join(dirpath, f)).read()
                    assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src, f


def test_params_mirror():
    from cardiax import params
    assert params.Params._fields == O.Params._fields
    assert params.MAXFLOAT == 1e6
    for k, v in O.PARAMSETS.items():
        assert tuple(getattr(params, "PARAMSET_" + k)) == tuple(v), k


def test_stimulus_builders_mirror():
    from cardiax import stimulus
    assert [d.value for d in stimulus.Direction] == [0, 1, 2, 3] and stimulus.Direction.NORTH == 0
    shape = (40, 60)
    p = stimulus.Protocol(0, 2, 1e9)
    for d in range(4):
        assert np.array_equal(stimulus.linear(shape, d, 0.2, 20.0, p).field.cpu().numpy(), O.linear(shape, d, 0.2, 20.0, p).field)
        assert np.array_equal(stimulus.triangular(shape, d, 33.0, 0.3, 20.0, p).field.cpu().numpy(),
                              O.triangular(shape, d, 33.0, 0.3, 20.0, p).field)
    assert np.array_equal(stimulus.rectangular(shape, (20, 30), (9, 5), 0.6, p).field.cpu().numpy(),
                          O.rectangular(shape, (20, 30), (9, 5), 0.6, p).field)
    r = stimulus.rectangular(shape, (20, 30), (9, 5), 0.6, p).field.cpu().numpy()
    assert r[15:24, 27:32].min() == np.float32(0.6) and np.count_nonzero(r) == 9 * 5  # int(20 -+ 4.5), int(30 -+ 2.5)
    with pytest.raises(ValueError):
        stimulus.linear(shape, 7, 0.2, 1.0, p)
    s = stimulus.linear(shape, stimulus.Direction.SOUTH, 0.5, 1.0, p)
    assert s.protocol is p and s._fields == ("protocol", "field")


def test_convert_mirror():
    from cardiax import convert
    assert convert.realsize_to_shape((12, 12), 0.01) == (1200, 1200)
    assert convert.ms_to_units(1000, 0.01) == 100000 and convert.cm_to_units(0.2, 0.01) == 20
    c = np.array([[0.0, 0.5], [0.25, 1.0]])
    out = convert.diffusivity_rescale(c, (1e-4, 1e-3))
    assert np.isclose(out.min(), 1e-4) and np.isclose(out.max(), 1e-3)
    assert convert.u_to_V(0.0) == -85 and convert.V_to_u(15.0) == 1.0


def test_api_surface_matches_reference_names():
    import inspect
    from cardiax import solve
    sig = lambda f: list(inspect.signature(f).parameters)
    assert solve.State._fields == ("v", "w", "u")
    assert sig(solve.step) == ["state", "t", "params", "diffusivity", "stimuli", "dx"]
    assert sig(solve.step_euler) == ["state", "t", "params", "diffusivity", "stimuli", "dt", "dx"]
    assert sig(solve._forward_euler) == ["state", "t", "t_end", "params", "diffusivity", "stimuli", "dt", "dx"]
    assert sig(solve.forward) == ["state", "checkpoints", "params", "diffusivity", "stimuli", "dt", "dx", "integrator", "plot_while"]
    assert sig(solve.forward_dimensional) == ["tissue_size", "final_time", "ms_step", "params", "diffusivity", "stimuli", "dt",
                                              "dx", "integrator", "plot_while"]
    assert sig(solve.gradient) == ["a", "axis"] and sig(solve.stimulate) == ["t", "X", "stimuli"]
    assert solve.TimeIntegrator.EULER == solve._forward_euler and callable(solve.TimeIntegrator.HEUN)
    assert inspect.signature(solve.forward).parameters["integrator"].default == solve.TimeIntegrator.EULER


def test_uniform_diffusivity_verdict_is_tied_to_the_tensor_object():
    """The verdict is cached for the caller's own tensor only (weak reference + version counter): an entry left behind
    under a recycled id / address must not be believed, and an in-place edit must be seen."""
    import weakref
    import torch
    from cardiax_b200 import solve
    solve._uniform_cache.clear()
    a = torch.full((8, 8), 1e-3)
    assert solve._is_uniform(a) is True
    assert solve._uniform_cache[id(a)][0]() is a
    a[3, 3] = 5e-4                       # in place: same object, new version
    assert solve._is_uniform(a) is False
    b = torch.full((8, 8), 1e-3)
    b[0, 0] = 2e-3
    other = torch.full((8, 8), 1e-3)
    solve._uniform_cache[id(b)] = (weakref.ref(other), b._version, True)    # a stale entry under b's id
    assert solve._is_uniform(b) is False
    dead = torch.full((8, 8), 1e-3)
    solve._uniform_cache[id(b)] = (weakref.ref(dead), b._version, True)
    del dead                              # ... and one whose tensor is gone
    assert solve._is_uniform(b) is False
