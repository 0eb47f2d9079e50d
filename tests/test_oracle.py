"""The oracle against everything the reference pins for this path and against itself."""
from fractions import Fraction

import numpy as np
import pytest

import oracle as O
from oracle import c_oracle as C
from tests import common

P3 = O.PARAMSETS["3"]


def test_stimulus_schedule_kat():
    """Reference tests/macro/stimulate_test.py:16-19 -- the only golden vector the reference has for this path."""
    A, B, Cc = O.Protocol(0, 2, 50), O.Protocol(10, 2, 50), O.Protocol(30, 2, 1000000)
    assert [t for t in range(300) if O.stimulus_active(t, A)] == [0, 1, 50, 51, 100, 101, 150, 151, 200, 201, 250, 251]
    assert [t for t in range(300) if O.stimulus_active(t, B)] == [10, 11, 60, 61, 110, 111, 160, 161, 210, 211, 260, 261]
    assert [t for t in range(300) if O.stimulus_active(t, Cc)] == [30, 31]
    shape = (80, 80)
    stim = [O.linear(shape, 0, 0.05, 1.0, A), O.triangular(shape, 3, 30, 0.5, 1.0, B), O.rectangular(shape, (50, 50), (1, 1), 1.0, Cc)]
    active = set([0, 1, 50, 51, 100, 101, 150, 151, 200, 201, 250, 251, 10, 11, 60, 61, 110, 111, 160, 161, 210, 211, 260,
                  261, 30, 31])
    X = np.zeros(shape, np.float32)
    for t in range(300):
        assert (np.count_nonzero(O.stimulate(t, X, stim)) != 0) == (t in active), t


def test_stimulus_unittest_expectation():
    """Reference tests/unittests/stimulus_test.py:12-23."""
    shape = (10, 10)
    s = O.stimulate(1, np.zeros(shape, np.float32), [O.linear(shape, 0, 0.5, 0.6, O.Protocol(0, 2, 1e9))])
    assert abs(float(np.mean(s[:5])) - 0.6) <= 1e-3
    assert np.all(s[5:] == 0)


def test_schedule_derived_form():
    """(3, 5, 20) fires at 3, 4, then [20..24], [40..44] ...; later stimuli override earlier ones."""
    p = O.Protocol(3, 5, 20)
    assert [t for t in range(50) if O.stimulus_active(t, p)] == [3, 4, 20, 21, 22, 23, 24, 40, 41, 42, 43, 44]
    a = O.Stimulus(O.Protocol(0, 2, 10), np.array([[1.0, 1.0, 0.0]], np.float32))
    b = O.Stimulus(O.Protocol(0, 2, 10), np.array([[0.0, -2.0, 0.0]], np.float32))
    out = O.stimulate(0, np.array([[7.0, 7.0, 7.0]], np.float32), [a, b])
    assert out.tolist() == [[1.0, -2.0, 7.0]]


def test_c_schedule_equals_numpy_schedule():
    L = C.lib()
    rng = np.random.default_rng(0)
    for _ in range(300):
        start, dur, per = int(rng.integers(0, 50)), int(rng.integers(1, 6)), float(rng.choice([7, 50, 400, 1e6, 1e9]))
        for t in range(0, 120):
            assert bool(L.fk_oracle_stim_active_f32(t, start, dur, per)) == O.stimulus_active(t, O.Protocol(start, dur, per))


def test_gradient_rows_exact_rationals():
    """solve.py:232-249 coefficients: rows 0,1 forward, rows n-2,n-1 backward, 4th-order centre (fp64 vs rationals)."""
    n = 11
    a = np.arange(n, dtype=np.float64) ** 3 / 7.0
    g = O.gradient(a, 0)
    fa = [Fraction(int(x) ** 3, 7) for x in range(n)]
    for i in range(n):
        if i < 2:
            e = Fraction(-11, 6) * fa[i] + 3 * fa[i + 1] - Fraction(3, 2) * fa[i + 2] + Fraction(1, 3) * fa[i + 3]
        elif i >= n - 2:
            e = Fraction(-1, 3) * fa[i - 3] + Fraction(3, 2) * fa[i - 2] - 3 * fa[i - 1] + Fraction(11, 6) * fa[i]
        else:
            e = Fraction(1, 12) * fa[i - 2] - Fraction(2, 3) * fa[i - 1] + Fraction(2, 3) * fa[i + 1] - Fraction(1, 12) * fa[i + 2]
        assert abs(g[i] - float(e)) < 1e-9
    # cubic is differentiated exactly by both the 3rd-order edges and the 4th-order centre
    assert np.allclose(g, 3 * np.arange(n) ** 2 / 7.0, atol=1e-9)
    b = np.random.default_rng(0).random((4, 6, 9))
    assert O.gradient(b, 2).shape == b.shape and O.gradient(b, -2).shape == b.shape


def test_dependence_radius():
    """d_u depends on u within a radius-4 plus and on D within a radius-2 plus."""
    shape = (31, 33)
    st, D = common.smooth_case(shape, 0)
    base = O.step(st, 5, P3, D, [], 0.01).u
    u2 = st.u.copy(); u2[15, 16] += 0.25
    diff = np.argwhere(O.step(st._replace(u=u2), 5, P3, D, [], 0.01).u != base)
    assert set(map(tuple, diff)) <= {(15 + k, 16) for k in range(-4, 5)} | {(15, 16 + k) for k in range(-4, 5)}
    assert (15 + 4, 16) in set(map(tuple, diff)) and (15, 16 - 4) in set(map(tuple, diff))
    D2 = D.copy(); D2[15, 16] *= 1.5
    diff = np.argwhere(O.step(st, 5, P3, D2, [], 0.01).u != base)
    assert set(map(tuple, diff)) <= {(15 + k, 16) for k in range(-2, 3)} | {(15, 16 + k) for k in range(-2, 3)}


def test_stimulus_replaces_current_not_voltage():
    """solve.py:46: d_u = del_u + stimulus where the (non-zero) stimulus field is active."""
    shape = (16, 16)
    st = O.init(shape)
    D = np.full(shape, 1e-3, np.float32)
    s = O.linear(shape, 0, 0.25, 20.0, O.Protocol(0, 2, 1e9))
    d = O.step(st, 0, P3, D, [s], 0.01)
    assert np.all(d.u[:4] == np.float32(20.0)) and np.all(d.u[4:] == 0)


@pytest.mark.parametrize("pset", ["3", "5", "2", "7", "1A", "10"])
def test_c_port_bit_identical_to_numpy(pset):
    st, D, stim = common.random_case((33, 47), seed=1, n_stim=3)
    a = O.forward_euler(st, 0, 25, O.PARAMSETS[pset], D, stim, 0.01, 0.01)
    b = C.forward_euler(st, 0, 25, O.PARAMSETS[pset], D, stim, 0.01, 0.01)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    a = O.forward_euler(st, 0, 10, O.PARAMSETS[pset], D, stim, 0.01, 0.01, dtype=np.float64)
    b = C.forward_euler(st, 0, 10, O.PARAMSETS[pset], D, stim, 0.01, 0.01, dtype=np.float64)
    for x, y in zip(a, b):
        assert np.allclose(x, y, rtol=0, atol=1e-13)


def test_segment_splitting_and_thread_count_invariance():
    st, D, stim = common.random_case((40, 40), seed=2)
    one = C.forward_euler(st, 0, 30, P3, D, stim, 0.01, 0.01)
    s = st
    for a, b in ((0, 7), (7, 8), (8, 30)):
        s = C.forward_euler(s, a, b, P3, D, stim, 0.01, 0.01)
    for x, y in zip(one, s):
        assert np.array_equal(x, y)
    C.set_threads(1)
    single = C.forward_euler(st, 0, 30, P3, D, stim, 0.01, 0.01)
    C.set_threads(0)
    for x, y in zip(one, single):
        assert np.array_equal(x, y)


def test_xla_tanh_restatement():
    """The restated XLA rational is a tanh: <= 4e-7 from libm over the clamp range, exactly x below 4e-4."""
    x = np.linspace(-12, 12, 400001).astype(np.float32)
    assert np.abs(O.tanh_xla_f32(x) - np.tanh(x.astype(np.float64))).max() < 4e-7
    small = np.float32([1e-4, -3.9e-4, 0.0])
    assert np.array_equal(O.tanh_xla_f32(small), small)
    assert O.tanh_xla_f32(np.float32([50.0, -1e11])).tolist() == O.tanh_xla_f32(np.float32([9.0, -9.0])).tolist()


def test_golden_wave_128_and_survey_smoke_values():
    """BASELINE config 1 shape.  u_max 0.2 -> 0.40 -> 0.90 -> 1.008 at steps 0/1/100/999."""
    import os
    shape = (128, 128)
    D = np.full(shape, 1e-3, np.float32)
    stim = [O.linear(shape, 0, 0.2, 20.0, O.Protocol(0, 2, 1e9))]
    s = O.init(shape)
    seen = {}
    for a, b in ((0, 1), (1, 2), (2, 101), (101, 1000)):
        s = C.forward_euler(s, a, b, P3, D, stim, 0.01, 0.01)
        seen[b - 1] = float(s.u.max())
    assert abs(seen[0] - 0.2) < 1e-6 and abs(seen[1] - 0.40208334) < 1e-6
    assert abs(seen[100] - 0.902864) < 1e-5 and abs(seen[999] - 1.0080984) < 1e-5
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "fk_128_wave_1000.npz"))
    assert np.array_equal(s.u[::4, ::4], g["u"]) and np.array_equal(s.v[:, 64], g["v_col"]) and np.array_equal(s.w[:, 64], g["w_col"])


def test_golden_scar_s1s2_fixture():
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "fk_64x96_scar_s1s2.npz"))
    stim = [O.Stimulus(O.Protocol(*g["proto%d" % i]), g["field%d" % i]) for i in range(2)]
    s = O.State(g["v0"], g["w0"], g["u0"])
    cps = g["checkpoints"]
    for i in range(len(cps) - 1):
        s = C.forward_euler(s, cps[i], cps[i + 1], O.Params(*g["params"]), g["D"], stim, float(g["dt"]), float(g["dx"]))
        for k, x in zip("vwu", s):
            assert np.array_equal(x, g["%s%d" % (k, i + 1)])


def test_fp32_drift_envelope_defines_the_tolerance():
    """fp32 oracle vs fp64 twin on the config-1 plane wave.  Two legitimate fp32 implementations (this oracle with
    XLA's rational tanh, the same with libm tanh, the fp64 twin rounded) sit 2e-6 .. 7e-6 apart up to 500 steps and
    1e-5 .. 4e-5 apart at 1e3 steps (the steep upstroke turns a tiny phase shift into a pointwise difference).
    The GPU tolerances are set against these measured figures: 2e-5 up to 200 steps, 1e-4 at 1e3 steps, and always
    |gpu - f64| <= 2 |oracle_f32 - f64| + 2e-6."""
    shape = (64, 64)
    D = np.full(shape, 1e-3, np.float32)
    stim = [O.linear(shape, 0, 0.2, 20.0, O.Protocol(0, 2, 1e9))]
    a = C.forward_euler(O.init(shape), 0, 1000, P3, D, stim, 0.01, 0.01)
    b = C.forward_euler(O.init(shape), 0, 1000, P3, D, stim, 0.01, 0.01, dtype=np.float64)
    drift = max(float(np.abs(x - y).max()) for x, y in zip(a, b))
    assert 1e-8 < drift < 1e-4, drift
    a2 = C.forward_euler(O.init(shape), 0, 200, P3, D, stim, 0.01, 0.01)
    b2 = C.forward_euler(O.init(shape), 0, 200, P3, D, stim, 0.01, 0.01, dtype=np.float64)
    assert max(float(np.abs(x - y).max()) for x, y in zip(a2, b2)) < 1e-5
    # the libm-tanh variant of the oracle is another legitimate fp32 implementation: same order of magnitude
    c = C.forward_euler(O.init(shape), 0, 1000, P3, D, stim, 0.01, 0.01, tanh="libm")
    assert max(float(np.abs(x - y).max()) for x, y in zip(a, c)) < 1e-4


def test_heun_restatement_is_second_order_and_matches_its_definition():
    """oracle.step_heun (cardiax/solve.py:73-85): the mean of the two slopes, both taken at the same counter; against the
    fp64 twin of a run with half the time step it converges with order 2 where Euler converges with order 1."""
    from tests import common
    st, D = common.smooth_case((24, 32), seed=1)
    P = O.PARAMSETS["3"]
    k1 = O.step(st, 0, P, D, [], 0.01)
    y1 = O.State(*[a + b * np.float32(0.01) for a, b in zip(st, k1)])
    k2 = O.step(y1, 0, P, D, [], 0.01)
    ref = O.State(*[a + (b + c) * np.float32(0.005) for a, b, c in zip(st, k1, k2)])
    got = O.step_heun(st, 0, P, D, [], 0.01, 0.01)
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)
    # order of convergence on u over a fixed horizon (fp64)
    def run(fn, dt, n):
        return fn(st, 0, n, P, D, [], dt, 0.01, dtype=np.float64)[2]
    exact = run(O.forward_heun, 0.00125, 64)
    e_h = [np.abs(run(O.forward_heun, dt, n) - exact).max() for dt, n in ((0.01, 8), (0.005, 16))]
    e_e = [np.abs(run(O.forward_euler, dt, n) - exact).max() for dt, n in ((0.01, 8), (0.005, 16))]
    assert e_h[0] / e_h[1] > 3.0 and 1.6 < e_e[0] / e_e[1] < 2.6 and e_h[0] < e_e[0]


# --------------------------------------------------------------------------- vectors frozen from the reference's own source
def _ref_fixture(name):
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", name))


def _fixture_stimuli(g):
    stim, i = [], 0
    while "field%d" % i in g.files:
        stim.append(O.Stimulus(O.Protocol(*[float(x) for x in g["proto%d" % i]]), g["field%d" % i]))
        i += 1
    return stim


@pytest.mark.parametrize("tanh,oracle_tanh", [("xla", "xla"), ("numpy", "libm")])
def test_oracle_equals_reference_vectors_integrators(tanh, oracle_tanh):
    """tests/golden/ref_fk_48x40.npz: `solve.forward` (Euler, Heun) and the int32-counter `_forward_euler` as computed by
    the UNMODIFIED reference source (make_reference_golden.py); needs no reference tree."""
    g = _ref_fixture("ref_fk_48x40.npz")
    P, stim, cps = O.Params(*g["params"]), _fixture_stimuli(g), g["checkpoints"]
    for name, fwd in (("euler", O.forward_euler), ("heun", O.forward_heun)):
        s = O.State(g["v0"], g["w0"], g["u0"])
        for i in range(len(cps) - 1):
            s = fwd(s, float(cps[i]), float(cps[i + 1]), P, g["D"], stim, float(g["dt"]), float(g["dx"]), tanh=oracle_tanh)
            for f, a in zip("vwu", s):
                assert np.array_equal(a, g["%s_%s_%s%d" % (name, tanh, f, i + 1)]), (name, f, i)
    s = O.forward_euler(O.State(g["v0"], g["w0"], g["u0"]), 0, 60, P, g["D"], stim, 0.01, 0.01, tanh=oracle_tanh, counter="i32")
    for f, a in zip("vwu", s):
        assert np.array_equal(a, g["euler_int_%s_%s" % (tanh, f)])
    if tanh == "xla":   # the C port (CPU baseline, long fixtures) reproduces the reference vectors as well
        s = O.State(g["v0"], g["w0"], g["u0"])
        for i in range(len(cps) - 1):
            s = C.forward_euler(s, float(cps[i]), float(cps[i + 1]), P, g["D"], stim, float(g["dt"]), float(g["dx"]))
            for f, a in zip("vwu", s):
                assert np.array_equal(a, g["euler_xla_%s%d" % (f, i + 1)]), ("C", f, i)


@pytest.mark.parametrize("tanh,oracle_tanh", [("xla", "xla"), ("numpy", "libm")])
def test_oracle_equals_reference_vectors_step_gradient_stimulate(tanh, oracle_tanh):
    """tests/golden/ref_fk_steps.npz: `solve.step` for all 16 parameter sets, N-D `gradient`, `stimulate` schedule."""
    g = _ref_fixture("ref_fk_steps.npz")
    st, stim = O.State(g["v0"], g["w0"], g["u0"]), _fixture_stimuli(g)
    for key, P in O.PARAMSETS.items():
        d = O.step(st, float(g["t"]), P, g["D"], stim, float(g["dx"]), tanh=oracle_tanh)
        for f, a in zip("vwu", d):
            assert np.array_equal(a, g["d%s_%s_%s" % (f, tanh, key)]), (key, f)
    for axis in range(4):
        assert np.array_equal(O.gradient(g["grad_in"], axis), g["grad_axis%d" % axis])
    for t in range(24):
        assert np.array_equal(O.stimulate(float(t), g["stimulate_in"], stim), g["stimulate_out"][t]), t
