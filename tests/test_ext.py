"""What surrounds the Euler loop: Dormand-Prince integrator (solve._forward_dormandprince / step_rk),
metrics.electrogram, the deepx.generate driver.  CPU tests run the product's element bodies and controller through the
emulation; `gpu` tests call the CUDA path through the Python API -> C ABI."""
import os

import numpy as np
import pytest
import torch

import oracle as O
from oracle import fk_oracle_ext as X
from tests import common
from tests.golden import make_dopri

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fk_dopri_24x28.npz")


# --------------------------------------------------------------------------- Dormand-Prince, CPU
def test_dopri_oracle_matches_its_golden_fixture():
    g = np.load(GOLD)
    st, D, stim = make_dopri.case()
    stats = {}
    ref = X.odeint_dopri5(st, g["ts"], O.PARAMSETS["3"], D, stim, 0.01, rtol=1e-5, atol=1e-5, stats=stats)
    for name, a in zip("vwu", ref):
        assert np.array_equal(a, g[name + "_loose"])
    assert [stats["attempts"], stats["accepted"], stats["rhs_evals"]] == list(g["stats_loose"])
    assert np.array_equal(ref.u[0], st.u)        # odeint returns y0 first (jnp.concatenate((y0[None], ys)))


def test_dopri_driver_and_bodies_bit_identical_to_the_oracle():
    """fk_ode.h (controller) + fk_aux.h (element bodies), compiled for the CPU, against the NumPy restatement."""
    from tests.emu import emu
    g = np.load(GOLD)
    st, D, stim = make_dopri.case()
    for name in ("tight", "loose"):
        tol = float(g["tol_" + name])
        (v, w, u), stats = emu.dopri5(st, g["ts"], O.PARAMSETS["3"], D, stim, 0.01, rtol=tol, atol=tol)
        assert np.array_equal(v, g["v_" + name]) and np.array_equal(w, g["w_" + name]) and np.array_equal(u, g["u_" + name])
        assert [stats["attempts"], stats["accepted"], stats["rhs_evals"]] == list(g["stats_" + name])
    # rejected attempts exist in this run (the controller's reject branch is exercised) and 6 evaluations per attempt + 2
    assert g["stats_tight"][0] > g["stats_tight"][1] and g["stats_tight"][2] == 6 * g["stats_tight"][0] + 2


def test_dopri_restatement_is_a_fifth_order_integrator():
    """Independent check of the restated algorithm: scipy's DOP853 at 1e-10 on the fp64 right-hand side.  The FK
    right-hand side is discontinuous (u >= V_c gates), so a few cells sit ~1e-3 apart whatever the integrator (scipy's
    own 1e-7 and 1e-10 runs differ by as much); the bulk must agree to fp32 round-off."""
    from scipy.integrate import solve_ivp
    shape = (24, 28)
    st, D = common.smooth_case(shape, 1)
    ts = np.array([0, 1.0, 2.5], np.float32)
    ref = X.odeint_dopri5(st, ts, O.PARAMSETS["3"], D, [], 0.01)
    n = shape[0] * shape[1]

    def f(t, y):
        s = O.State(*[y[i * n:(i + 1) * n].reshape(shape) for i in range(3)])
        return np.concatenate([x.reshape(-1) for x in O.step(s, t, O.PARAMSETS["3"], D.astype(np.float64), [], 0.01,
                                                             dtype=np.float64, tanh="libm")])
    y0 = np.concatenate([np.asarray(x, np.float64).reshape(-1) for x in st])
    sol = solve_ivp(f, (0, 2.5), y0, method="DOP853", rtol=1e-10, atol=1e-12, t_eval=[1.0, 2.5])
    for j in (0, 1):
        for i in range(3):
            d = np.abs(sol.y[i * n:(i + 1) * n, j].reshape(shape) - ref[i][j + 1])
            assert np.median(d) < 2e-6 and d.max() < 5e-3, (i, j, float(np.median(d)), float(d.max()))


def test_dopri_tableau_equals_an_independent_statement_of_dormand_prince():
    """Known-answer check of the restated Butcher tableau: nodes, stage weights and the 5th-order solution weights of
    jax.experimental.ode (alpha, beta, c_sol) are Dormand-Prince's, as SciPy states them independently (RK45.C / A / B).
    (The error weights differ by construction: jax / TensorFlow use Shampine's embedded 4th-order solution, SciPy the
    original one -- so c_error and c_mid stay anchored on the integrator-level check above only.)"""
    from scipy.integrate._ivp.rk import RK45
    from oracle import fk_oracle_ext as X
    f32 = np.float32
    assert [f32(c) for c in RK45.C[1:]] == X._ALPHA[:5] and X._ALPHA[5] == f32(1.0)
    for s_, row in enumerate(RK45.A[1:]):                      # stages 2 .. 6
        assert [f32(a) for a in row[:s_ + 1]] == X._BETA[s_][:s_ + 1], s_
        assert all(b == 0 for b in X._BETA[s_][s_ + 1:])
    assert [f32(b) for b in RK45.B] == X._C_SOL[:6] and X._C_SOL[6] == 0
    assert X._BETA[5] == X._C_SOL                              # first-same-as-last: the 7th stage is the next step's first
    # the embedded solution has order 4: its weights sum to 1 and satisfy the first three order conditions as well
    c4 = np.array(X._C_SOL, np.float64) - np.array(X._C_ERR, np.float64)
    nodes = np.array([0.0] + [float(a) for a in X._ALPHA[:6]])
    assert abs(c4.sum() - 1) < 1e-6 and abs(c4 @ nodes - 0.5) < 1e-6 and abs(c4 @ nodes ** 2 - 1 / 3) < 1e-6 \
        and abs(c4 @ nodes ** 3 - 0.25) < 1e-6


def test_dopri_controller_scalars():
    """optimal_step_size of jax.experimental.ode: grow by at most 10x, shrink by at most 5x, safety 0.9."""
    f = X._optimal_step_size
    assert f(np.float32(0.1), np.float32(0.0)) == np.float32(0.1) * np.float32(10)
    assert np.isclose(f(np.float32(0.1), np.float32(1.0)), 0.09, rtol=1e-6)                # ratio 1 -> dt * 0.9
    assert np.isclose(f(np.float32(0.1), np.float32(1e12)), 0.02, rtol=1e-6)               # clipped at 1 / dfactor
    assert np.isclose(f(np.float32(0.1), np.float32(1e-30)), 1.0, rtol=1e-6)               # clipped at ifactor


def test_electrogram_body_matches_the_oracle():
    from tests.emu import emu
    rng = np.random.default_rng(0)
    x = rng.random((5, 20, 20)).astype(np.float32)
    for point in ((3, 7), (0, 0), (9.5, 2.5)):
        got, ref = emu.electrogram(x, point), X.electrogram(x, point)
        assert got.shape == (5,) and np.allclose(got, ref, rtol=2e-7, atol=0)
    one = np.zeros((20, 20), np.float32)
    one[4, 6] = 2.0                                   # x[i=4][j=6] * sqrt((6 - 3)^2 + (4 - 0)^2) = 2 * 5
    assert emu.electrogram(one, (3, 0)) == 10.0 and X.electrogram(one, (3, 0)) == 10.0


def test_plot_calls_degrade_gracefully_without_matplotlib():
    """solve.forward(plot_while=True) / generate.sequence(plot_while=True) call cardiax.plot (solve.py:179-186, 218-220):
    visualisation is out of scope, but the calls must not break a run on a box without matplotlib."""
    import warnings
    import cardiax
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        st = O.init((8, 8))
        for out in (cardiax.plot.plot_state(st), cardiax.plot.plot_diffusivity(np.ones((8, 8))),
                    cardiax.plot.plot_stimuli([O.linear((8, 8), 0, 0.5, 1.0, O.Protocol(0, 2, 10))]), cardiax.plot.plot_stimuli([])):
            assert out is None or len(out) == 2


# --------------------------------------------------------------------------- generate (host logic)
def test_random_generators_follow_the_reference_ranges():
    from cardiax_b200 import generate
    shape = (120, 120)
    kinds = set()
    for seed in range(40):
        s = generate.random_stimulus(seed, shape, min_start=7, max_start=8)
        f = np.asarray(s.field.cpu())
        assert f.shape == shape and f.dtype == np.float32
        assert int(np.asarray(s.protocol.start).reshape(-1)[0]) == 7 and s.protocol.duration == 2
        assert 400 <= int(np.asarray(s.protocol.period).reshape(-1)[0]) < 10 ** 9
        assert np.abs(f).max() <= 20.0 * 1.6              # cubic-spline rotation overshoots a little
        kinds.add(bool(((f != 0) & (f != 20.0)).any()))
        s2 = generate.random_stimulus(seed, shape, min_start=7, max_start=8)
        assert np.array_equal(f, np.asarray(s2.field.cpu()))     # seeded
    assert kinds == {True, False}                                # rotated (non-binary) and axis-aligned masks both drawn
    D = generate.random_diffusivity(3, shape)
    assert D.shape == shape and D.dtype == np.float32 and np.isclose(D.min(), 1e-4) and np.isclose(D.max(), 1e-3)
    p = generate.random_protocol(0)
    assert 0 <= int(p.start[0]) < 1000 and p.start.shape == (1,)


def test_ensemble_sharding_covers_every_seed_once():
    from cardiax_b200 import generate
    for n, world in ((1024, 8), (1024, 1), (5, 2), (3, 8), (0, 4), (17, 4)):
        parts = [generate.shard(range(n), r, world) for r in range(world)]
        assert sum(parts, []) == list(range(n))                 # contiguous, in order, nothing twice
        assert max(len(p) for p in parts) == -(-n // world)     # ceil(n / world) per rank, the tail gets the rest
    assert [len(generate.shard(range(1024), r, 8)) for r in range(8)] == [128] * 8     # BASELINE config 4


# --------------------------------------------------------------------------- GPU
def _gpu_stimuli(stim):
    from cardiax_b200 import stimulus
    return [stimulus.Stimulus(stimulus.Protocol(*s.protocol), torch.as_tensor(s.field).cuda()) for s in stim]


@pytest.mark.gpu
def test_dopri_gpu_exact_bitwise_vs_oracle_fixture():
    from cardiax_b200 import _lib, options, solve
    g = np.load(GOLD)
    st, D, stim = make_dopri.case()
    gst = _gpu_stimuli(stim)
    state = solve.State(*[torch.as_tensor(x).cuda() for x in st])
    Dg = torch.as_tensor(D).cuda()
    old = (options.numerics, options.ode_rtol, options.ode_atol)
    try:
        options.numerics = "exact"
        for name in ("tight", "loose"):
            options.ode_rtol = options.ode_atol = float(g["tol_" + name])
            before = _lib.lib().fk_launch_count()
            out = solve._forward_dormandprince(state, g["ts"], O.PARAMSETS["3"], Dg, gst, 0.01, 0.01)
            assert _lib.lib().fk_launch_count() - before > 8 * int(g["stats_" + name][0])
            assert out.u.shape == (4, 24, 28)
            for nm, a in zip("vwu", out):
                assert np.array_equal(a.cpu().numpy(), g[nm + "_" + name]), (name, nm)
            s = solve.last_ode_stats
            assert [s["attempts"], s["accepted"], s["rhs_evals"]] == list(g["stats_" + name])
        # fast numerics: same algorithm, FMA arithmetic -- a different but equally valid adaptive trajectory.  The bar
        # is MEASURED like the Euler path's: the gap between two valid oracles (XLA's rational tanh vs libm's) at the
        # same tolerance; fast numerics must stay within twice that envelope, per snapshot.  (It is wide at t = 1:
        # the stimulus current acts for two whole time units there, u reaches 33 and the gates are discontinuous.)
        options.numerics = "fast"
        out = solve._forward_dormandprince(state, g["ts"], O.PARAMSETS["3"], Dg, gst, 0.01, 0.01)
        alt = X.odeint_dopri5(st, g["ts"], O.PARAMSETS["3"], D, stim, 0.01, rtol=1e-5, atol=1e-5, tanh="libm")
        for nm, a, b in zip("vwu", out, alt):
            ref = g[nm + "_loose"]
            for j in range(len(g["ts"])):
                d, env = np.abs(a[j].cpu().numpy() - ref[j]), np.abs(b[j] - ref[j])
                assert np.median(d) <= 1e-5, (nm, j, float(np.median(d)))
                assert np.quantile(d, 0.99) <= 2 * np.quantile(env, 0.99) + 1e-5, (nm, j, float(np.quantile(d, 0.99)))
                assert d.max() <= 2 * env.max() + 1e-2, (nm, j, float(d.max()), float(env.max()))
    finally:
        options.numerics, options.ode_rtol, options.ode_atol = old


@pytest.mark.gpu
def test_forward_with_the_dormandprince_integrator():
    """solve.forward(..., integrator=TimeIntegrator.DORMANDPRINCE) returns odeint's stacked State (solve.py:189-195);
    step_rk is the same call (solve.py:88-89)."""
    from cardiax_b200 import options, solve
    shape = (40, 36)
    st, D = common.smooth_case(shape, 2)
    state = solve.State(*[torch.as_tensor(x).cuda() for x in st])
    Dg = torch.as_tensor(D).cuda()
    old = (options.numerics, options.ode_rtol, options.ode_atol, options.verbose)
    try:
        options.numerics, options.ode_rtol, options.ode_atol, options.verbose = "exact", 1e-5, 1e-5, False
        cps = np.arange(0, 4, 1)
        out = solve.forward(state, cps, O.PARAMSETS["5"], Dg, [], 0.01, 0.01, integrator=solve.TimeIntegrator.DORMANDPRINCE,
                            plot_while=True)
        assert isinstance(out, solve.State) and tuple(out.u.shape) == (4,) + shape
        ref = X.odeint_dopri5(st, cps, O.PARAMSETS["5"], D, [], 0.01, rtol=1e-5, atol=1e-5)
        for a, b in zip(out, ref):
            assert np.array_equal(a.cpu().numpy(), b)
        rk = solve.step_rk(state, cps, O.PARAMSETS["5"], Dg, [], 0.01, 0.01)
        assert all(torch.equal(a, b) for a, b in zip(rk, out))
        # a batch is integrated as one system: two copies of the same tissue stay identical to each other
        b2 = solve.State(*[torch.stack([x, x]) for x in state])
        ob = solve._forward_dormandprince(b2, cps, O.PARAMSETS["5"], Dg, [], 0.01, 0.01)
        assert tuple(ob.u.shape) == (4, 2) + shape and torch.equal(ob.u[:, 0], ob.u[:, 1])
    finally:
        options.numerics, options.ode_rtol, options.ode_atol, options.verbose = old


@pytest.mark.gpu
def test_electrogram_gpu():
    from cardiax_b200 import metrics
    rng = np.random.default_rng(0)
    x = rng.random((2, 5, 64, 64)).astype(np.float32)
    for point in ((3, 7), (31.5, 0.5)):
        got = metrics.electrogram(torch.as_tensor(x).cuda(), point).cpu().numpy()
        assert got.shape == (2, 5) and np.allclose(got, X.electrogram(x, point), rtol=2e-7, atol=0)
    with pytest.raises(ValueError):
        metrics.electrogram(torch.zeros((4, 6), device="cuda"), (0, 0))
    assert metrics.adp(None, 0.9) is None and metrics.spiral_centres(None) is None


@pytest.mark.gpu
def test_generate_sequence_and_ensemble(tmp_path):
    from cardiax_b200 import generate, io, options, params, solve
    old = options.verbose
    options.verbose = False
    try:
        shape, out = (96, 96), (24, 24)
        path = os.path.join(tmp_path, "seq.hdf5")
        generate.random_sequence(5, params.PARAMSET_3, path, shape=shape, n_stimuli=1, start=0, stop=1.0, step=0.25,
                                 dt=0.01, dx=0.01, reshape=out, plot_while=False)
        f = io._open(path, "r")
        assert f["states"].shape == (4, 3, *out) and f["field"].shape == (1, *out) and f["diffusivity"].shape == out
        # same inputs through the plain API give the same snapshots
        stimuli, D = generate._random_inputs(5, params.PARAMSET_3, shape, 1, 1.0, 0.01)
        s = solve.init(shape)
        for i in range(3):
            s = solve._forward_euler(s, i * 25, (i + 1) * 25, params.PARAMSET_3, D, stimuli, 0.01, 0.01)
            assert np.allclose(f["states"][i], io.imresize(torch.stack(tuple(s)), out).cpu().numpy(), atol=1e-6)
        assert np.abs(f["states"][2][2]).max() > 0.1      # the first stimulus (start = 1) excited the tissue
        # ensemble: 5 seeds over 2 "ranks", batched; each member equals its own random_sequence
        pattern = os.path.join(tmp_path, "ens_%d.hdf5")
        mine0 = generate.ensemble(range(5), params.PARAMSET_3, pattern, shape=shape, n_stimuli=1, stop=1.0, step=0.25,
                                  reshape=out, rank=0, world_size=2, chunk=2)
        mine1 = generate.ensemble(range(5), params.PARAMSET_3, pattern, shape=shape, n_stimuli=1, stop=1.0, step=0.25,
                                  reshape=out, rank=1, world_size=2, chunk=2)
        assert mine0 == [0, 1, 2] and mine1 == [3, 4]
        generate.random_sequence(3, params.PARAMSET_3, os.path.join(tmp_path, "single.hdf5"), shape=shape, n_stimuli=1,
                                 stop=1.0, step=0.25, reshape=out, plot_while=False)
        a, b = io._open(pattern % 3, "r"), io._open(os.path.join(tmp_path, "single.hdf5"), "r")
        assert np.array_equal(a["states"][:3], b["states"][:3]) and np.array_equal(a["field"][:], b["field"][:])
    finally:
        options.verbose = old
