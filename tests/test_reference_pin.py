"""Pins ``oracle/`` to the reference's OWN SOURCE: ``/root/reference/cardiax/{solve,stimulus,params,convert}.py`` are
imported unmodified, by path, on the NumPy stand-in for jax of ``tests/golden/ref_shim.py``, and every function on the
hot path must return the same bits as the oracle's restatement.

Skipped where the reference tree is absent (the GPU box): there the committed vectors that
``tests/golden/make_reference_golden.py`` froze from the same runs take over (``tests/test_oracle.py``,
``tests/test_gpu_parity.py``).
"""
import numpy as np
import pytest

import oracle as O
from tests import common
from tests.golden import ref_shim as S

pytestmark = pytest.mark.skipif(not S.available(), reason="reference sources not mounted at %s" % S.REFERENCE_ROOT)

PAIRS = [("1A", "PARAMSET_1A"), ("1B", "PARAMSET_1B"), ("1C", "PARAMSET_1C"), ("1D", "PARAMSET_1D"), ("1E", "PARAMSET_1E"),
         ("2", "PARAMSET_2"), ("3", "PARAMSET_3"), ("4A", "PARAMSET_4A"), ("4B", "PARAMSET_4B"), ("4C", "PARAMSET_4C"),
         ("5", "PARAMSET_5"), ("6", "PARAMSET_6"), ("7", "PARAMSET_7"), ("8", "PARAMSET_8"), ("9", "PARAMSET_9"),
         ("10", "PARAMSET_10")]
TANH = [("numpy", "libm"), ("xla", "xla")]     # (stand-in's switch, oracle's switch): the same function on both sides


@pytest.fixture(scope="module")
def ref():
    return S.load_reference()


def rstim(ref, stimuli):
    return [ref.stimulus.Stimulus(ref.stimulus.Protocol(*s.protocol), s.field) for s in stimuli]


def same(a, b):
    a, b = S.to_numpy(a), tuple(b)
    assert all(x.dtype == np.float32 and np.isfinite(x).all() for x in a)
    return all(np.array_equal(x, y) for x, y in zip(a, b))


def test_reference_files_are_the_pinned_ones():
    assert S.file_hashes() == S.REFERENCE_SHA256


def test_loading_the_reference_leaves_no_fake_jax_behind(ref):
    import sys
    assert "jax" not in sys.modules or not hasattr(sys.modules["jax"], "tree_multimap") or sys.modules["jax"].__name__ != "jax" \
        or getattr(sys.modules["jax"], "__file__", None) is not None
    assert not any(k.startswith("_fk_reference_cardiax") for k in sys.modules)


def test_params_table_is_the_reference_table(ref):
    from cardiax_b200 import params as ours
    assert ref.params.Params._fields == O.Params._fields == ours.Params._fields
    assert ref.params.MAXFLOAT == ours.MAXFLOAT
    names = [n for n in dir(ref.params) if n.startswith("PARAMSET_")]
    assert sorted(names) == sorted(n for _, n in PAIRS)
    for key, name in PAIRS:
        assert tuple(getattr(ref.params, name)) == tuple(O.PARAMSETS[key]) == tuple(getattr(ours, name)), name


def test_init(ref):
    a = S.to_numpy(ref.solve.init((5, 7)))
    assert same(a, O.init((5, 7))) and type(a)._fields == ("v", "w", "u")


@pytest.mark.parametrize("key,name", PAIRS)
@pytest.mark.parametrize("shim_tanh,oracle_tanh", TANH)
def test_step_all_paramsets(ref, key, name, shim_tanh, oracle_tanh):
    """solve.py:26-65 on white-noise and smooth states, heterogeneous D, three stimuli, at active and idle counters."""
    P, RP = O.PARAMSETS[key], getattr(ref.params, name)
    st, D, stim = common.random_case((21, 26), seed=3, n_stim=3)
    sm, Ds = common.smooth_case((19, 23), seed=4)
    with ref.tanh(shim_tanh):
        for t in (0, 1.0, 2, 4.0, 11):
            assert same(ref.solve.step(ref.solve.State(*st), t, RP, D, rstim(ref, stim), 0.01),
                        O.step(st, t, P, D, stim, 0.01, tanh=oracle_tanh)), (key, t)
        assert same(ref.solve.step(ref.solve.State(*sm), 7.0, RP, Ds, [], 0.025), O.step(sm, 7.0, P, Ds, [], 0.025, tanh=oracle_tanh))


@pytest.mark.parametrize("shim_tanh,oracle_tanh", TANH)
def test_step_euler_and_forward_euler_float_counter(ref, shim_tanh, oracle_tanh):
    """solve.py:68-70, 92-100: the float32 counter `solve.forward` produces; periodic stimuli fire several times."""
    st, D, stim = common.random_case((24, 28), seed=0, n_stim=3)
    P, RP = O.PARAMSETS["3"], ref.params.PARAMSET_3
    with ref.tanh(shim_tanh):
        assert same(ref.solve.step_euler(ref.solve.State(*st), 3.0, RP, D, rstim(ref, stim), 0.01, 0.01),
                    O.step_euler(st, 3.0, P, D, stim, 0.01, 0.01, tanh=oracle_tanh))
        a = ref.solve._forward_euler(ref.solve.State(*st), 0.0, 20.0, RP, D, rstim(ref, stim), 0.01, 0.01)
        assert same(a, O.forward_euler(st, 0.0, 20.0, P, D, stim, 0.01, 0.01, tanh=oracle_tanh))
        sm, Ds = common.smooth_case((24, 28), seed=6)     # white noise blows up after ~40 steps; a smooth field does not
        a = ref.solve._forward_euler(ref.solve.State(*sm), 0.0, 60.0, RP, Ds, rstim(ref, stim), 0.01, 0.01)
        assert same(a, O.forward_euler(sm, 0.0, 60.0, P, Ds, stim, 0.01, 0.01, tanh=oracle_tanh))
        assert same(a, O.forward_euler(sm, 0.0, 60.0, P, Ds, stim, 0.01, 0.01, tanh=oracle_tanh, counter="f32"))
        # a segment that starts in the middle of a period, other dt / dx, homogeneous D
        st2, D2, stim2 = common.random_case((17, 33), seed=5, n_stim=2, hetero=False)
        assert same(ref.solve._forward_euler(ref.solve.State(*st2), 5.0, 31.0, ref.params.PARAMSET_5, D2, rstim(ref, stim2), 0.02, 0.025),
                    O.forward_euler(st2, 5.0, 31.0, O.PARAMSETS["5"], D2, stim2, 0.02, 0.025, tanh=oracle_tanh))


def test_c_port_matches_the_reference_source(ref):
    """The C + OpenMP port (the CPU baseline and the generator of the long fixtures) against the reference itself."""
    from oracle import c_oracle
    st, D = common.smooth_case((40, 36), seed=2)
    stim = [O.linear((40, 36), 0, 0.2, 20.0, O.Protocol(0, 2, 1e9)), O.rectangular((40, 36), (20, 18), (8, 8), 20.0, O.Protocol(25, 2, 40))]
    for shim_tanh, oracle_tanh in TANH[1:]:   # libm's tanhf is not NumPy's SIMD tanh: only the "xla" pair is bit-comparable
        with ref.tanh(shim_tanh):
            a = ref.solve._forward_euler(ref.solve.State(*st), 0.0, 100.0, ref.params.PARAMSET_3, D, rstim(ref, stim), 0.01, 0.01)
        assert same(a, c_oracle.forward_euler(st, 0, 100, O.PARAMSETS["3"], D, stim, 0.01, 0.01, tanh=oracle_tanh))


def test_forward_euler_int_counter_and_array_protocols(ref):
    """deepx/generate.py:24-27, 187-196: int32 counter, `start` / `period` as shape-(1,) int32 arrays, duration a Python
    int, one period a Python float (1e9) -- every combination jax's promotion distinguishes."""
    shape = (16, 20)
    st, D, _ = common.random_case(shape, seed=7, n_stim=0)
    f = np.zeros(shape, np.float32)
    f[:4] = 20.0
    g = np.zeros(shape, np.float32)
    g[:, -5:] = -3.0
    protos = [(np.array([2], np.int32), 2, np.array([9], np.int32)), (4, 3, 1e9), (np.array([1], np.int32), 2.0, 6)]
    stim = [O.Stimulus(O.Protocol(*p), fld) for p, fld in zip(protos, (f, g, f * 0.5))]
    a = ref.solve._forward_euler(ref.solve.State(*st), 0, 30, ref.params.PARAMSET_3, D, rstim(ref, stim), 0.01, 0.01)
    assert same(a, O.forward_euler(st, 0, 30, O.PARAMSETS["3"], D, stim, 0.01, 0.01, tanh="libm", counter="i32"))


def test_schedule_typed_semantics_beyond_2_pow_24(ref):
    """The int32 path stays exact where float32 rounds: periods and counters above 2^24 (missing 6)."""
    X = np.zeros((1, 2), np.float32)
    fld = np.ones((1, 2), np.float32)
    cases = [
        (np.array([3], np.int32), 2, np.array([16777259], np.int32)),        # odd period > 2^24: float32 rounds it to ...260
        (0, 2, 33554467), (5, 4, np.array([123456789], np.int32)), (7, 2, 1e9), (7.0, 2, 400), (0, 2, 50)]
    for start, dur, per in cases:
        proto = O.Protocol(start, dur, per)
        s0, p0 = int(np.asarray(start).reshape(-1)[0]), int(np.asarray(per).reshape(-1)[0])
        ts = sorted(set([0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10] + [s0 + k * p0 + d for k in (1, 2) for d in range(-4, 4)]))
        for t in ts:
            if t >= 2 ** 31 - 8:
                continue
            for tt in (int(t), np.int32(t), float(t)):
                got = bool(np.asarray(ref.solve.stimulate(tt, X, [ref.stimulus.Stimulus(ref.stimulus.Protocol(*proto), fld)])).any())
                assert got == O.stimulus_active_typed(tt, proto), (start, dur, per, tt)
    # and the two typings DO part above 2^24, which is why the C ABI carries the types (include/fk.h, FkStimulus v2)
    pi, pf = O.Protocol(3, 2, 16777259), O.Protocol(3.0, 2.0, 16777259.0)
    apart = [t for t in range(3 + 16777259 - 40, 3 + 16777259 + 40)
             if O.stimulus_active_typed(int(t), pi) != O.stimulus_active_typed(float(t), pf)]
    assert apart == [16777264, 16777265]
    for t in apart:
        for tt, proto in ((int(t), pi), (float(t), pf)):
            got = bool(np.asarray(ref.solve.stimulate(tt, X, [ref.stimulus.Stimulus(ref.stimulus.Protocol(*proto), fld)])).any())
            assert got == O.stimulus_active_typed(tt, proto)


def test_stimulate_schedule_kat_on_the_reference(ref):
    """The reference's own known-answer vector (tests/macro/stimulate_test.py:16-19), run through its own stimulate."""
    shape = (80, 80)
    A, B, C = O.Protocol(0, 2, 50), O.Protocol(10, 2, 50), O.Protocol(30, 2, 1000000)
    stim = [O.linear(shape, 0, 0.05, 1.0, A), O.triangular(shape, 3, 30, 0.5, 1.0, B), O.rectangular(shape, (50, 50), (1, 1), 1.0, C)]
    active = {0, 1, 50, 51, 100, 101, 150, 151, 200, 201, 250, 251, 10, 11, 60, 61, 110, 111, 160, 161, 210, 211, 260, 261, 30, 31}
    X = np.zeros(shape, np.float32)
    rs = rstim(ref, stim)
    for t in range(300):
        out = S.to_numpy(ref.solve.stimulate(t, X, rs))
        assert (np.count_nonzero(out) != 0) == (t in active), t
        assert np.array_equal(out, O.stimulate(t, X, stim)), t
    # overriding order, negative amplitudes, zero cells (solve.py:269-271)
    X = np.random.default_rng(0).standard_normal(shape).astype(np.float32)
    stim2 = [O.Stimulus(O.Protocol(0, 2, 10), stim[0].field), O.Stimulus(O.Protocol(0, 2, 10), -2.0 * stim[1].field)]
    assert np.array_equal(S.to_numpy(ref.solve.stimulate(1.0, X, rstim(ref, stim2))), O.stimulate(1.0, X, stim2))


@pytest.mark.parametrize("shim_tanh,oracle_tanh", TANH)
def test_heun(ref, shim_tanh, oracle_tanh):
    """solve.py:73-85, 103-111."""
    st, D, stim = common.random_case((20, 24), seed=1, n_stim=3)
    P, RP = O.PARAMSETS["4A"], ref.params.PARAMSET_4A
    with ref.tanh(shim_tanh):
        assert same(ref.solve.step_heun(ref.solve.State(*st), 3.0, RP, D, rstim(ref, stim), 0.01, 0.01),
                    O.step_heun(st, 3.0, P, D, stim, 0.01, 0.01, tanh=oracle_tanh))
        assert same(ref.solve._forward_heun(ref.solve.State(*st), 0.0, 25.0, RP, D, rstim(ref, stim), 0.01, 0.01),
                    O.forward_heun(st, 0.0, 25.0, P, D, stim, 0.01, 0.01, tanh=oracle_tanh))
        assert same(ref.solve._forward_heun(ref.solve.State(*st), 0, 12, RP, D, rstim(ref, stim), 0.03, 0.02),
                    O.forward_heun(st, 0, 12, P, D, stim, 0.03, 0.02, tanh=oracle_tanh, counter="i32"))


def test_gradient_nd_any_axis(ref):
    """solve.py:225-254, incl. the 5-D / negative-axis use of deepx/optimise.py:27-30."""
    rng = np.random.default_rng(0)
    a = rng.standard_normal((5, 6, 7, 5, 9)).astype(np.float32)
    for axis in (0, 1, 2, 3, 4, -1, -2):
        assert np.array_equal(S.to_numpy(ref.solve.gradient(a, axis)), O.gradient(a, axis)), axis
    b = rng.standard_normal((5, 41)).astype(np.float32)
    for axis in (0, 1):
        assert np.array_equal(S.to_numpy(ref.solve.gradient(b, axis)), O.gradient(b, axis))


def test_forward_checkpoint_loop(ref, capsys):
    """solve.py:168-222: list of States at checkpoints[1:], float(checkpoint) counters."""
    st, D, stim = common.random_case((18, 22), seed=9, n_stim=2)
    cps = np.arange(0, 40, 10)
    got = ref.solve.forward(ref.solve.State(*st), cps, ref.params.PARAMSET_3, D, rstim(ref, stim), 0.01, 0.01)
    capsys.readouterr()
    assert len(got) == len(cps) - 1
    s = st
    for i in range(len(cps) - 1):
        s = O.forward_euler(s, float(cps[i]), float(cps[i + 1]), O.PARAMSETS["3"], D, stim, 0.01, 0.01, tanh="libm")
        assert same(got[i], s)


def test_forward_dimensional_shape_assertion(ref, capsys):
    """solve.py:133-165: AssertionError on a diffusivity / stimulus shape mismatch; checkpoints = arange(0, stop, step)."""
    P = ref.params.PARAMSET_3
    D = np.full((12, 16), 1e-3, np.float32)
    stim = [ref.stimulus.linear((12, 16), ref.stimulus.Direction.NORTH, 0.3, 20.0, ref.stimulus.Protocol(0, 2, 1e9))]
    with pytest.raises(AssertionError):
        ref.solve.forward_dimensional((0.12, 0.17), 0.4, 0.1, P, D, stim, 0.01, 0.01)
    got = ref.solve.forward_dimensional((0.121, 0.161), 0.401, 0.101, P, D, stim, 0.01, 0.01)
    capsys.readouterr()
    cps = np.arange(0, int(0.401 / 0.01), int(0.101 / 0.01))
    ost = [O.Stimulus(O.Protocol(0, 2, 1e9), S.to_numpy(stim[0].field))]
    s = O.init((12, 16))
    assert len(got) == len(cps) - 1
    for i in range(len(cps) - 1):
        s = O.forward_euler(s, float(cps[i]), float(cps[i + 1]), O.PARAMSETS["3"], D, ost, 0.01, 0.01, tanh="libm")
        assert same(got[i], s)


def test_mask_builders(ref):
    """stimulus.py:31-140 against the oracle's and the product's builders (same fields, same ValueError)."""
    from cardiax_b200 import stimulus as ours
    RS, proto = ref.stimulus, (0, 2, 1e9)
    for shape in ((40, 56), (64, 48)):
        for centre, size in (((20, 30), (10, 6)), ((5, 5), (20, 20)), ((33, 12), (7, 9))):
            a = S.to_numpy(RS.rectangular(shape, centre, size, 0.6, RS.Protocol(*proto)).field)
            assert np.array_equal(a, O.rectangular(shape, centre, size, 0.6, O.Protocol(*proto)).field)
            assert np.array_equal(a, ours.rectangular(shape, centre, size, 0.6, ours.Protocol(*proto)).field.cpu().numpy())
        for d in range(4):
            for cov in (0.05, 0.2, 0.5):
                a = S.to_numpy(RS.linear(shape, RS.Direction(d), cov, 20.0, RS.Protocol(*proto)).field)
                assert a.dtype == np.float32
                assert np.array_equal(a, O.linear(shape, d, cov, 20.0, O.Protocol(*proto)).field)
                assert np.array_equal(a, ours.linear(shape, ours.Direction(d), cov, 20.0, ours.Protocol(*proto)).field.cpu().numpy())
                for angle in (10.0, 30, 100.0):
                    b = np.asarray(RS.triangular(shape, RS.Direction(d), angle, cov, 20.0, RS.Protocol(*proto)).field)
                    assert b.dtype == np.float32
                    assert np.array_equal(b, O.triangular(shape, d, angle, cov, 20.0, O.Protocol(*proto)).field)
                    assert np.array_equal(b, ours.triangular(shape, ours.Direction(d), angle, cov, 20.0,
                                                             ours.Protocol(*proto)).field.cpu().numpy())
    for mod in (RS, ours):
        with pytest.raises(ValueError):
            mod.linear((8, 8), 7, 0.5, 1.0, mod.Protocol(*proto))
    assert [int(x) for x in RS.Direction] == [int(x) for x in ours.Direction] and RS.Direction.__members__.keys() == ours.Direction.__members__.keys()
    assert RS.Protocol._fields == ours.Protocol._fields and RS.Stimulus._fields == ours.Stimulus._fields


def test_convert(ref):
    """convert.py:23-63 against the product's module (pure host arithmetic)."""
    from cardiax_b200 import convert as ours
    RC = ref.convert
    for fn, args in (("realsize_to_shape", ((12.3, 4.56), 0.01)), ("shape_to_realsize", ((1200, 1150), 0.01)), ("cm_to_units", (3.7, 0.01)),
                     ("units_to_cm", (370, 0.01)), ("ms_to_units", (410.5, 0.01)), ("units_to_ms", (41050, 0.01)), ("u_to_V", (0.37,)),
                     ("V_to_u", (-40.0,)), ("diffusivity_to_units", (0.05, 0.01))):
        assert getattr(RC, fn)(*args) == getattr(ours, fn)(*args), fn
    c = np.random.default_rng(0).random((9, 11)).astype(np.float32)
    assert np.array_equal(np.asarray(RC.diffusivity_rescale(c, (1e-4, 1e-3))), np.asarray(ours.diffusivity_rescale(c, (1e-4, 1e-3))))
    p = dict(ref.params.PARAMSET_3._asdict())
    assert RC.params_to_units(dict(p), 0.01, 0.02) == ours.params_to_units(dict(p), 0.01, 0.02)
    st = [dict(start=1.0, duration=0.02, period=4.0)]
    assert RC.stimuli_to_units([dict(s) for s in st], 0.01, 0.01) == ours.stimuli_to_units([dict(s) for s in st], 0.01, 0.01)
