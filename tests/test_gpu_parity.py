"""GPU parity tests: the CUDA path, called through the reference-facing Python API (which goes
through the C ABI of libfk.so), against the CPU oracle on the same seeded inputs.

Bars
  numerics="exact": BIT-EXACT against oracle.fk_oracle (np.array_equal; -0.0 == +0.0).
  numerics="fast":  max-abs(u, v, w) <= 2e-5 up to 200 steps (TOL_FAST), <= 1e-4 at 1e3 steps (TOL_FAST_1K), and
                    never further from the fp64 twin than twice the fp32 oracle is (+ 2e-6).  These figures are set
                    against the measured distance between legitimate fp32 implementations of the reference
                    (tests/test_oracle.py::test_fp32_drift_envelope_defines_the_tolerance).
"""
import numpy as np
import pytest
import torch

import oracle as O
from oracle import c_oracle as C
from tests import common

pytestmark = pytest.mark.gpu

TOL_FAST = 2e-5
TOL_FAST_1K = 1e-4
P3 = O.PARAMSETS["3"]


@pytest.fixture(autouse=True)
def _reset_options():
    from cardiax_b200 import options
    saved = {k: getattr(options, k) for k in ("numerics", "steps_per_launch", "kernel", "cta_threads", "rows_per_cta",
                                              "tiles", "cells_per_thread", "edge_tile", "maps_global")}
    options.verbose = False
    yield
    for k, v in saved.items():
        setattr(options, k, v)


def run_gpu(st, t0, t1, params, D, stim, dt=0.01, dx=0.01, **opts):
    from cardiax_b200 import options, solve, stimulus
    for k, v in opts.items():
        setattr(options, k, v)
    gstim = [stimulus.Stimulus(stimulus.Protocol(*s.protocol), torch.as_tensor(s.field).cuda()) for s in stim]
    out = solve._forward_euler(solve.State(*[torch.as_tensor(x).cuda() for x in st]), t0, t1, params,
                               torch.as_tensor(D).cuda(), gstim, dt, dx)
    torch.cuda.synchronize()
    return [x.cpu().numpy() for x in out]


def assert_exact(got, ref, what=""):
    for name, a, b in zip("vwu", got, ref):
        assert not np.isnan(a).any(), "%s %s has NaN" % (what, name)
        assert np.array_equal(a, b), "%s %s: max abs diff %g at %d cells" % (
            what, name, np.abs(a - b).max(), int((a != b).sum()))


@pytest.mark.parametrize("shape,T,kernel,extra", [
    ((37, 53), 1, 1, {}), ((37, 53), 2, 1, {}), ((37, 53), 3, 1, {}), ((5, 7), 2, 1, {}), ((3, 3), 1, 1, {}),
    ((128, 128), 2, 0, {}), ((72, 160), 1, 2, dict(cta_threads=32, rows_per_cta=16)),
    ((72, 160), 2, 2, dict(cta_threads=32, rows_per_cta=16)), ((96, 288), 3, 2, dict(cta_threads=64, rows_per_cta=20)),
    ((200, 1200), 2, 2, {}), ((256, 256), 2, 0, {}), ((130, 516), 4, 2, dict(cta_threads=128)),
    ((128, 128), 0, 3, {}), ((12, 16), 0, 3, {}), ((3, 4), 0, 3, {}), ((512, 512), 0, 0, {}), ((9, 132), 0, 3, {}),
])
def test_exact_bitwise_vs_oracle(shape, T, kernel, extra):
    st, D, stim = common.random_case(shape, seed=3, n_stim=3)
    ref = C.forward_euler(st, 0, 9, P3, D, stim, 0.01, 0.01)
    got = run_gpu(st, 0, 9, P3, D, stim, numerics="exact", steps_per_launch=T, kernel=kernel, **extra)
    assert_exact(got, ref, "shape %s T %d kernel %d" % (shape, T, kernel))


@pytest.mark.parametrize("shape,nsteps,tiles,threads,nc", [
    ((12, 16), 9, (0, 0), 0, 0), ((3, 4), 5, (0, 0), 0, 0), ((64, 96), 9, (2, 3), 0, 4), ((64, 96), 9, (4, 1), 256, 2),
    ((128, 128), 9, (0, 0), 0, 0), ((100, 200), 9, (5, 5), 0, 0), ((256, 256), 40, (0, 0), 0, 0), ((512, 512), 9, (0, 0), 0, 0),
    ((512, 512), 12, (16, 8), 512, 4), ((33, 72), 6, (4, 9), 128, 1), ((64, 96), 1, (2, 3), 0, 0), ((1024, 1024), 6, (0, 0), 0, 0),
    ((256, 256), 30, (8, 8), 0, 2), ((128, 128), 20, (16, 1), 64, 4), ((200, 120), 25, (3, 3), 96, 1)])
def test_resident_kernel_exact_bitwise_vs_oracle(shape, nsteps, tiles, threads, nc):
    """The resident kernel (whole call in one cooperative launch; halos exchanged between co-resident CTAs through
    tagged 8-byte mailbox records in L2): bit-identical to the oracle for any tile grid, CTA size and cells per thread
    (threads < groups: several rounds per step)."""
    from cardiax_b200 import _lib
    st, D, stim = common.random_case(shape, seed=4, n_stim=3)
    ref = C.forward_euler(st, 0, nsteps, P3, D, stim, 0.01, 0.01)
    before = _lib.lib().fk_launch_count()
    got = run_gpu(st, 0, nsteps, P3, D, stim, numerics="exact", kernel=4, tiles=tiles, cta_threads=threads,
                  cells_per_thread=nc)
    assert _lib.lib().fk_launch_count() - before == 2   # D_x/D_y maps + ONE resident launch
    assert_exact(got, ref, "resident %s tiles %s" % (shape, tiles))


@pytest.mark.parametrize("shape,nsteps,tiles,threads,nc", [
    ((12, 16), 9, (0, 0), 0, 0), ((64, 96), 9, (2, 3), 0, 4), ((64, 96), 9, (4, 1), 256, 2), ((128, 128), 40, (0, 0), 0, 0),
    ((128, 128), 9, (4, 4), 0, 4), ((128, 128), 9, (2, 8), 128, 1), ((100, 200), 9, (3, 5), 0, 0), ((256, 256), 40, (0, 0), 0, 0),
    ((64, 64), 33, (0, 0), 0, 0), ((320, 320), 7, (4, 4), 0, 0), ((64, 96), 1, (2, 3), 0, 0), ((200, 120), 25, (3, 3), 96, 1)])
def test_cluster_kernel_exact_bitwise_vs_oracle(shape, nsteps, tiles, threads, nc):
    """The resident kernel's cluster form (a tissue of <= 16 tiles = ONE thread-block cluster; ring cells stored into the
    neighbouring CTAs' halos through distributed shared memory, one cluster barrier per Euler step): bit-identical to the
    oracle for any tile grid, CTA size and cells per thread."""
    from cardiax_b200 import _lib
    st, D, stim = common.random_case(shape, seed=4, n_stim=3)
    ref = C.forward_euler(st, 0, nsteps, P3, D, stim, 0.01, 0.01)
    before = _lib.lib().fk_launch_count()
    got = run_gpu(st, 0, nsteps, P3, D, stim, numerics="exact", kernel=5, tiles=tiles, cta_threads=threads,
                  cells_per_thread=nc)
    assert _lib.lib().fk_launch_count() - before == 2   # D_x/D_y maps + ONE launch
    assert _lib.last_kernel() == "fk_cluster_kernel"
    assert_exact(got, ref, "cluster %s tiles %s" % (shape, tiles))


def test_cluster_kernel_fast_matches_the_other_kernels_and_takes_any_batch():
    """Fast numerics: cluster == mailbox-resident == streaming bit for bit over 300 steps with stimuli; a batch of 40
    tissues (more clusters than the device holds at once: they are independent, scheduled in waves) against single runs."""
    from cardiax_b200 import _lib, options, solve, stimulus
    st, D = common.smooth_case((64, 64), seed=8)
    _, _, stim = common.random_case((64, 64), seed=8, n_stim=3)
    a = run_gpu(st, 0, 300, P3, D, stim, numerics="fast")
    assert _lib.last_kernel() == "fk_cluster_kernel"        # the default for a single tissue up to 80 x 80
    b = run_gpu(st, 0, 300, P3, D, stim, numerics="fast", kernel=4)
    assert _lib.last_kernel() == "fk_resident_kernel"
    c = run_gpu(st, 0, 300, P3, D, stim, numerics="fast", kernel=3)
    assert_exact(a, b, "cluster vs resident")
    assert_exact(a, c, "cluster vs wide")
    shape, batch = (64, 64), 40
    cases = [common.random_case(shape, seed=30 + b, n_stim=2) for b in range(batch)]
    stb = [np.stack([cs[0][k] for cs in cases]) for k in range(3)]
    Db = np.stack([cs[1] for cs in cases])
    options.numerics, options.kernel = "exact", 5
    gst = [[stimulus.Stimulus(stimulus.Protocol(*s.protocol), torch.as_tensor(s.field).cuda()) for s in cs[2]] for cs in cases]
    out = solve._forward_euler(solve.State(*[torch.as_tensor(x).cuda() for x in stb]), 0, 11, P3, torch.as_tensor(Db).cuda(), gst, 0.01, 0.01)
    torch.cuda.synchronize()
    assert _lib.last_kernel() == "fk_cluster_kernel"
    for b in range(0, batch, 7):
        ref = C.forward_euler(cases[b][0], 0, 11, P3, cases[b][1], cases[b][2], 0.01, 0.01)
        assert_exact([x[b].cpu().numpy() for x in out], ref, "tissue %d" % b)


@pytest.mark.parametrize("shape,nsteps,tiles,edge,mg", [((1200, 1200), 5, (0, 0), (0, 0), 0), ((96, 160), 9, (5, 6), (12, 4), 1),
                                                        ((200, 120), 12, (3, 3), (0, 0), 1), ((512, 512), 8, (0, 0), (0, 0), 0)])
def test_resident_kernel_maps_in_l2_and_uneven_edge_tiles(shape, nsteps, tiles, edge, mg):
    """The reference's data-generation tissue (1200 x 1200, heterogeneous D: deepx/generate.py:92) keeps u, v, w in
    shared memory and reads D, D_x, D_y from L2; smaller tiles at the tissue's edges.  Bit-identical to the oracle."""
    from cardiax_b200 import _lib
    st, D, stim = common.random_case(shape, seed=6, n_stim=2)
    ref = C.forward_euler(st, 0, nsteps, P3, D, stim, 0.01, 0.01)
    got = run_gpu(st, 0, nsteps, P3, D, stim, numerics="exact", kernel=4, tiles=tiles, edge_tile=edge, maps_global=mg)
    assert _lib.last_kernel() == "fk_resident_kernel"
    if shape == (1200, 1200):
        assert _lib.last_plan()["maps_in_l2"] == 1
    assert_exact(got, ref, "resident %s" % (shape,))


def test_resident_kernel_is_the_default_for_small_tissues_and_matches_the_other_kernels():
    """Fast numerics are tiling- and kernel-independent: resident == wide == streaming, bit for bit, over 300 steps of
    a smooth excitable field with stimuli firing on the way (BASELINE config 1/2 sized tissues)."""
    from cardiax_b200 import _lib
    st, D = common.smooth_case((256, 256), seed=8)
    _, _, stim = common.random_case((256, 256), seed=8, n_stim=3)
    before = _lib.lib().fk_launch_count()
    a = run_gpu(st, 0, 300, P3, D, stim, numerics="fast")
    assert _lib.lib().fk_launch_count() - before == 2
    b = run_gpu(st, 0, 300, P3, D, stim, numerics="fast", kernel=3)
    c = run_gpu(st, 0, 300, P3, D, stim, numerics="fast", kernel=2, steps_per_launch=2)
    assert_exact(a, b, "resident vs wide")
    assert_exact(a, c, "resident vs streaming")


def test_resident_kernel_batched_and_repeated_calls():
    """Several tissues in one resident launch, then the same call again (flags are re-zeroed per call)."""
    shape, batch = (64, 64), 6
    cases = [common.random_case(shape, seed=30 + b, n_stim=2) for b in range(batch)]
    st = [np.stack([c[0][k] for c in cases]) for k in range(3)]
    D = np.stack([c[1] for c in cases])
    from cardiax_b200 import options, solve, stimulus
    options.numerics, options.kernel = "exact", 4
    gst = [[stimulus.Stimulus(stimulus.Protocol(*s.protocol), torch.as_tensor(s.field).cuda()) for s in c[2]] for c in cases]
    state = solve.State(*[torch.as_tensor(x).cuda() for x in st])
    for _ in range(2):
        out = solve._forward_euler(state, 0, 11, P3, torch.as_tensor(D).cuda(), gst, 0.01, 0.01)
        torch.cuda.synchronize()
        for b in range(batch):
            ref = C.forward_euler(cases[b][0], 0, 11, P3, cases[b][1], cases[b][2], 0.01, 0.01)
            assert_exact([x[b].cpu().numpy() for x in out], ref, "tissue %d" % b)


@pytest.mark.parametrize("shape,n", [((48, 64), 5), ((37, 53), 3), ((128, 128), 4)])
def test_heun_exact_bitwise_vs_oracle(shape, n):
    """solve._forward_heun / step_heun (cardiax/solve.py:73-85, 103-111): the device-side Heun loop (two right-hand sides
    at the same counter + fused stage kernels) is bit-identical to the oracle's literal restatement."""
    from cardiax_b200 import options, solve, stimulus
    st, D, stim = common.random_case(shape, seed=9, n_stim=3)
    ref = O.forward_heun(st, 0, n, P3, D, stim, 0.01, 0.01)
    options.numerics = "exact"
    gst = [stimulus.Stimulus(stimulus.Protocol(*s.protocol), torch.as_tensor(s.field).cuda()) for s in stim]
    gstate = solve.State(*[torch.as_tensor(x).cuda() for x in st])
    got = solve._forward_heun(gstate, 0, n, P3, torch.as_tensor(D).cuda(), gst, 0.01, 0.01)
    assert_exact([x.cpu().numpy() for x in got], ref, "heun %s" % (shape,))
    one = solve.step_heun(gstate, 0, P3, torch.as_tensor(D).cuda(), gst, 0.01, 0.01)
    assert_exact([x.cpu().numpy() for x in one], O.step_heun(st, 0, P3, D, stim, 0.01, 0.01), "step_heun")
    # through the checkpoint loop with the integrator selected like the reference does (TimeIntegrator.HEUN)
    states = solve.forward(gstate, [0, 2, n], P3, torch.as_tensor(D).cuda(), gst, 0.01, 0.01, solve.TimeIntegrator.HEUN)
    assert_exact([x.cpu().numpy() for x in states[-1]], ref, "forward(HEUN)")
    options.numerics = "fast"
    fast = solve._forward_heun(gstate, 0, n, P3, torch.as_tensor(D).cuda(), gst, 0.01, 0.01)
    for a, b in zip(fast, ref):
        assert float(np.abs(a.cpu().numpy() - b).max()) <= TOL_FAST * max(1.0, float(np.abs(b).max()))


def test_heun_fast_through_the_euler_kernels_matches_exact_heun():
    """Fast numerics run a Heun step as y + (E(E(y)) - y) / 2 with E = one Euler step of the streaming (large tissue)
    or wide (small tissue) kernel, both stages at the same counter; exact numerics run the reference's literal formula
    in one tile-kernel launch per step (bit-identical to the oracle, test above).  Same tolerance as fast Euler."""
    from cardiax_b200 import _lib, options, solve, stimulus
    for shape, kernel in (((1024, 1056), "fk_stream_kernel"), ((96, 128), "fk_wide_kernel")):
        st, D = common.smooth_case(shape, seed=3)
        stim = [O.linear(shape, 0, 0.2, 20.0, O.Protocol(2, 2, 50))]
        gst = [stimulus.Stimulus(stimulus.Protocol(*s.protocol), torch.as_tensor(s.field).cuda()) for s in stim]
        gstate = solve.State(*[torch.as_tensor(x).cuda() for x in st])
        Dg = torch.as_tensor(D).cuda()
        options.numerics = "exact"
        ref = solve._forward_heun(gstate, 0, 6, P3, Dg, gst, 0.01, 0.01)
        assert _lib.last_kernel() == "fk_tile_kernel"
        options.numerics = "fast"
        fast = solve._forward_heun(gstate, 0, 6, P3, Dg, gst, 0.01, 0.01)
        assert _lib.last_kernel() == kernel
        for a, b in zip(fast, ref):
            assert float((a - b).abs().max()) <= TOL_FAST * max(1.0, float(b.abs().max()))
        assert float((fast.u - gstate.u).abs().max()) > 1e-3      # the stimulus at t = 2, 3 acted
        # the closing pass folded into the last launch's store (default) == the separate combine pass (kernel = 1)
        before = _lib.lib().fk_launch_count()
        solve._forward_heun(gstate, 0, 6, P3, Dg, gst, 0.01, 0.01)
        folded_launches = _lib.lib().fk_launch_count() - before
        options.kernel = 1
        before = _lib.lib().fk_launch_count()
        unfolded = solve._forward_heun(gstate, 0, 6, P3, Dg, gst, 0.01, 0.01)
        assert _lib.lib().fk_launch_count() - before == folded_launches + 6
        options.kernel = 0
        assert all(torch.equal(a, b) for a, b in zip(fast, unfolded))


@pytest.mark.parametrize("pset", sorted(O.PARAMSETS))
def test_exact_all_paramsets(pset):
    st, D, stim = common.random_case((64, 96), seed=5)
    ref = C.forward_euler(st, 0, 6, O.PARAMSETS[pset], D, stim, 0.01, 0.01)
    got = run_gpu(st, 0, 6, O.PARAMSETS[pset], D, stim, numerics="exact", kernel=2, cta_threads=32, rows_per_cta=24)
    assert_exact(got, ref, "paramset " + pset)


def test_exact_division_selftest_all_paramsets():
    """The 3-instruction correctly rounded division by constants equals __fdiv_rn for every significand (device)."""
    import ctypes
    from cardiax_b200 import _lib
    L = _lib.lib()
    for k, p in O.PARAMSETS.items():
        for dx in (0.01, 0.03, 0.005, 0.1):
            P = _lib.FkParams(*[np.float32(x) for x in p])
            bad = ctypes.c_longlong(-1)
            _lib.check(L.fk_check_exact_division(ctypes.byref(P), np.float32(dx), ctypes.byref(bad), None))
            assert bad.value == 0, (k, dx, bad.value)


def test_exact_safe_division_option_gives_the_same_bits():
    from cardiax_b200 import options
    st, D, stim = common.random_case((64, 160), seed=13)
    a = run_gpu(st, 0, 9, P3, D, stim, numerics="exact", kernel=2, cta_threads=32)
    options.safe_division = True
    try:
        b = run_gpu(st, 0, 9, P3, D, stim, numerics="exact", kernel=2, cta_threads=32)
    finally:
        options.safe_division = False
    assert_exact(a, b, "safe division")


def test_exact_uniform_diffusivity_fast_path():
    st, _, stim = common.random_case((96, 640), seed=7)
    D = np.full((96, 640), 1e-3, np.float32)
    ref = C.forward_euler(st, 0, 8, P3, D, stim, 0.01, 0.01)
    got = run_gpu(st, 0, 8, P3, D, stim, numerics="exact", kernel=2)
    assert_exact(got, ref, "uniform D")


def test_exact_wave_1000_steps_128():
    """BASELINE config 1: 128 x 128, PARAMSET_3, D = 1e-3, NORTH stripe, 1e3 steps -- bit-exact to the end."""
    shape = (128, 128)
    st = O.init(shape)
    D = np.full(shape, 1e-3, np.float32)
    stim = [O.linear(shape, 0, 0.2, 20.0, O.Protocol(0, 2, 1e9))]
    ref = C.forward_euler(st, 0, 1000, P3, D, stim, 0.01, 0.01)
    got = run_gpu(st, 0, 1000, P3, D, stim, numerics="exact")
    assert_exact(got, ref, "config 1")
    assert abs(float(got[2].max()) - 1.0080984) < 1e-6  # survey smoke value


def test_fast_within_tolerance_and_f64_envelope():
    shape = (128, 128)
    st = O.init(shape)
    D = np.full(shape, 1e-3, np.float32)
    stim = [O.linear(shape, 0, 0.2, 20.0, O.Protocol(0, 2, 1e9))]
    ref32 = C.forward_euler(st, 0, 1000, P3, D, stim, 0.01, 0.01)
    ref64 = C.forward_euler(st, 0, 1000, P3, D, stim, 0.01, 0.01, dtype=np.float64)
    got = run_gpu(st, 0, 1000, P3, D, stim, numerics="fast")
    for name, g, a, b in zip("vwu", got, ref32, ref64):
        err = np.abs(g - a).max()
        assert err <= TOL_FAST_1K, "%s: fast vs oracle_f32 %g" % (name, err)
        assert np.abs(g - b).max() <= 2 * np.abs(a - b).max() + 2e-6, name


@pytest.mark.parametrize("kernel,extra", [(1, {}), (2, dict(cta_threads=64, rows_per_cta=32))])
def test_fast_hetero_scar_stimuli(kernel, extra):
    (st, D) = common.smooth_case((160, 320), seed=2)
    _, _, stim = common.random_case((160, 320), seed=2, n_stim=3)
    ref = C.forward_euler(st, 0, 200, P3, D, stim, 0.01, 0.01)
    ref64 = C.forward_euler(st, 0, 200, P3, D, stim, 0.01, 0.01, dtype=np.float64)
    got = run_gpu(st, 0, 200, P3, D, stim, numerics="fast", kernel=kernel, **extra)
    for name, g, a, b in zip("vwu", got, ref, ref64):
        assert np.abs(g - a).max() <= TOL_FAST, name
        assert np.abs(g - b).max() <= 2 * np.abs(a - b).max() + 2e-6, name


def test_fast_is_tiling_independent():
    """Every cell's arithmetic is the same whichever kernel/tiling/T produced it -> identical bits."""
    (st, D) = common.smooth_case((136, 520), seed=4)
    _, _, stim = common.random_case((136, 520), seed=4, n_stim=2)
    a = run_gpu(st, 0, 12, P3, D, stim, numerics="fast", kernel=1, steps_per_launch=1)
    for kw in (dict(kernel=1, steps_per_launch=3), dict(kernel=2, steps_per_launch=2, cta_threads=64, rows_per_cta=24),
               dict(kernel=2, steps_per_launch=4, cta_threads=128), dict(kernel=2, steps_per_launch=1),
               dict(kernel=3, steps_per_launch=0, cta_threads=0, rows_per_cta=0)):
        b = run_gpu(st, 0, 12, P3, D, stim, numerics="fast", **kw)
        assert_exact(b, a, str(kw))


def test_segment_splitting_is_exact():
    """forward()'s checkpoint semantics: [0, 37) in one call == [0, 10) + [10, 23) + [23, 37)."""
    st, D, stim = common.random_case((96, 256), seed=9)
    one = run_gpu(st, 0, 37, P3, D, stim, numerics="fast")
    s = st
    for a, b in ((0, 10), (10, 23), (23, 37)):
        s = O.State(*run_gpu(s, a, b, P3, D, stim, numerics="fast"))
    assert_exact(list(s), one, "segments")


def test_batch_of_tissues_matches_single_runs():
    from cardiax_b200 import options, solve, stimulus
    options.numerics = "exact"
    B, shape = 3, (64, 128)
    cases = [common.random_case(shape, seed=20 + b, n_stim=2) for b in range(B)]
    # different schedules per tissue
    per = [[O.Stimulus(O.Protocol(b, 2, 5 + b), s.field) for s in cases[b][2]] for b in range(B)]
    v = torch.as_tensor(np.stack([c[0].v for c in cases])).cuda()
    w = torch.as_tensor(np.stack([c[0].w for c in cases])).cuda()
    u = torch.as_tensor(np.stack([c[0].u for c in cases])).cuda()
    D = torch.as_tensor(np.stack([c[1] for c in cases])).cuda()
    gst = [[stimulus.Stimulus(stimulus.Protocol(*s.protocol), torch.as_tensor(s.field).cuda()) for s in p] for p in per]
    out = solve._forward_euler(solve.State(v, w, u), 0, 11, P3, D, gst, 0.01, 0.01)
    for b in range(B):
        ref = C.forward_euler(cases[b][0], 0, 11, P3, cases[b][1], per[b], 0.01, 0.01)
        assert_exact([x[b].cpu().numpy() for x in out], ref, "tissue %d" % b)


def test_step_gradient_stimulate_api():
    from cardiax_b200 import options, solve, stimulus
    options.numerics = "exact"
    st, D, stim = common.random_case((40, 56), seed=11, n_stim=3)
    gstim = [stimulus.Stimulus(stimulus.Protocol(*s.protocol), torch.as_tensor(s.field).cuda()) for s in stim]
    gs = solve.State(*[torch.as_tensor(x).cuda() for x in st])
    for t in (0, 1, 3, 5):
        d = solve.step(gs, t, P3, torch.as_tensor(D).cuda(), gstim, 0.01)
        ref = O.step(st, t, P3, D, stim, 0.01)
        assert_exact([x.cpu().numpy() for x in d], ref, "step t=%d" % t)
        x = np.random.default_rng(t).random((40, 56), dtype=np.float32)
        s = solve.stimulate(t, torch.as_tensor(x).cuda(), gstim).cpu().numpy()
        assert np.array_equal(s, O.stimulate(t, x, stim))
    a = np.random.default_rng(0).random((3, 7, 9, 11), dtype=np.float32)
    for axis in (0, 1, 2, 3, -1, -2):
        if a.shape[axis] >= 5:
            g = solve.gradient(torch.as_tensor(a).cuda(), axis).cpu().numpy()
            assert np.array_equal(g, O.gradient(a, axis)), axis
    # step_euler == one Euler step of the oracle
    e = solve.step_euler(gs, 3, P3, torch.as_tensor(D).cuda(), gstim, 0.01, 0.01)
    assert_exact([x.cpu().numpy() for x in e], O.step_euler(st, 3, P3, D, stim, 0.01, 0.01), "step_euler")


def test_stimulus_schedule_kat_on_device():
    """tests/macro/stimulate_test.py:16-19 of the reference, through solve.stimulate on the GPU."""
    from cardiax_b200 import solve, stimulus
    shape = (80, 80)
    A = stimulus.linear(shape, stimulus.Direction.NORTH, 0.05, 1.0, stimulus.Protocol(0, 2, 50))
    B = stimulus.triangular(shape, stimulus.Direction.WEST, 30, 0.5, 1.0, stimulus.Protocol(10, 2, 50))
    Cc = stimulus.rectangular(shape, (50, 50), (1, 1), 1.0, stimulus.Protocol(30, 2, 1000000))
    active = {0, 1, 50, 51, 100, 101, 150, 151, 200, 201, 250, 251, 10, 11, 60, 61, 110, 111, 160, 161, 210, 211, 260,
              261, 30, 31}
    X = torch.zeros(shape, device="cuda")
    for t in range(300):
        nz = int((solve.stimulate(t, X, [A, B, Cc]) != 0).sum().item())
        assert (nz != 0) == (t in active), t


def test_forward_checkpoints_and_golden_fixture():
    import os
    from cardiax_b200 import options, solve, stimulus
    options.numerics = "exact"
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "fk_64x96_scar_s1s2.npz"))
    stim = [stimulus.Stimulus(stimulus.Protocol(*g["proto%d" % i]), torch.as_tensor(g["field%d" % i]).cuda())
            for i in range(2)]
    params = O.Params(*g["params"])
    states = solve.forward(solve.State(*[torch.as_tensor(g[k + "0"]).cuda() for k in "vwu"]), g["checkpoints"], params,
                           torch.as_tensor(g["D"]).cuda(), stim, float(g["dt"]), float(g["dx"]))
    assert len(states) == len(g["checkpoints"]) - 1
    for i, s in enumerate(states):
        assert_exact([x.cpu().numpy() for x in s], [g["%s%d" % (k, i + 1)] for k in "vwu"], "checkpoint %d" % i)


def test_large_field_properties_4096():
    """BASELINE config 3 size: no oracle run; size-independent properties instead."""
    from cardiax_b200 import options, solve
    options.numerics = "fast"
    H = W = 4096
    g = torch.Generator(device="cuda").manual_seed(0)
    u = torch.zeros(H, W, device="cuda")
    idx = torch.randint(0, H - 64, (40, 2), generator=g, device="cuda").cpu().numpy()
    for r, c in idx:
        u[r:r + 48, c:c + 48] = 1.0
    st = solve.State(torch.ones(H, W, device="cuda"), torch.ones(H, W, device="cuda"), u)
    D = torch.full((H, W), 1e-3, device="cuda")
    P5 = O.PARAMSETS["5"]
    a = solve._forward_euler(st, 0, 24, P5, D, [], 0.01, 0.01)
    # (1) segment splitting, (2) T and kernel independence
    b = solve._forward_euler(solve._forward_euler(st, 0, 7, P5, D, [], 0.01, 0.01), 7, 24, P5, D, [], 0.01, 0.01)
    options.kernel, options.steps_per_launch = 1, 1
    c = solve._forward_euler(st, 0, 24, P5, D, [], 0.01, 0.01)
    for x, y, z in zip(a, b, c):
        assert torch.equal(x, y) and torch.equal(x, z)
        assert torch.isfinite(x).all()
    # (3) a corner crop evolves exactly like the oracle's run of a tissue that contains its dependency cone
    n = 8
    sub = 96
    crop = O.State(*[x[:sub + 4 * n, :sub + 4 * n].cpu().numpy() for x in st])
    Dc = np.full(crop.u.shape, 1e-3, np.float32)
    options.numerics, options.kernel, options.steps_per_launch = "exact", 0, 0
    e = solve._forward_euler(st, 0, n, P5, D, [], 0.01, 0.01)
    ref = C.forward_euler(crop, 0, n, P5, Dc, [], 0.01, 0.01)
    for x, r in zip(e, ref):
        assert np.array_equal(x[:sub, :sub].cpu().numpy(), r[:sub, :sub])


@pytest.mark.parametrize("uniform", [True, False])
def test_large_field_four_corners_4096_exact(uniform):
    """BASELINE config 3 size at the launch geometry of the bench (one wave of 9 strips x 49 balanced row chunks: the
    chunks at the physical top / bottom edge are shorter, the edge strips' CTAs come first in block order): the crops at
    all four corners evolve exactly like the oracle's runs of tissues that contain their dependency cones -- the first /
    last strip, the first / last chunk (general body at the top, tail at the bottom) and their intersections."""
    from cardiax_b200 import _lib, options, solve
    H = W = 4096
    n, sub = 8, 96
    m = sub + 4 * n
    rng = np.random.default_rng(7)
    u = torch.zeros(H, W, device="cuda")
    v = torch.ones(H, W, device="cuda")
    w = torch.ones(H, W, device="cuda")
    D = torch.full((H, W), 1e-3, device="cuda")
    for r0, c0 in ((0, 0), (0, W - m), (H - m, 0), (H - m, W - m)):   # something to diffuse in every corner
        u[r0:r0 + m, c0:c0 + m] = torch.as_tensor(rng.random((m, m), dtype=np.float32)).cuda()
        v[r0:r0 + m, c0:c0 + m] = torch.as_tensor(rng.random((m, m), dtype=np.float32)).cuda()
        w[r0:r0 + m, c0:c0 + m] = torch.as_tensor(rng.random((m, m), dtype=np.float32)).cuda()
        if not uniform:
            D[r0:r0 + m, c0:c0 + m] = torch.as_tensor((rng.random((m, m), dtype=np.float32) * 9e-4 + 1e-4).astype(np.float32)).cuda()
    st = solve.State(v, w, u)
    P5 = O.PARAMSETS["5"]
    options.numerics, options.kernel, options.steps_per_launch = "exact", 2, 2
    try:
        e = solve._forward_euler(st, 0, n, P5, D, [], 0.01, 0.01)
        plan = _lib.last_plan()
    finally:
        options.numerics, options.kernel, options.steps_per_launch = "fast", 0, 0
    assert _lib.last_kernel() == "fk_stream_kernel" and plan["strips"] >= 3 and plan["row_chunks"] >= 4
    for r0, c0, rs, cs in ((0, 0, slice(0, sub), slice(0, sub)), (0, W - m, slice(0, sub), slice(m - sub, m)),
                           (H - m, 0, slice(m - sub, m), slice(0, sub)), (H - m, W - m, slice(m - sub, m), slice(m - sub, m))):
        crop = O.State(*[x[r0:r0 + m, c0:c0 + m].cpu().numpy() for x in st])
        Dc = D[r0:r0 + m, c0:c0 + m].cpu().numpy()
        ref = C.forward_euler(crop, 0, n, P5, Dc, [], 0.01, 0.01)
        for name, x, r in zip("vwu", e, ref):
            got = x[r0:r0 + m, c0:c0 + m].cpu().numpy()
            assert np.array_equal(got[rs, cs], r[rs, cs]), (name, r0, c0, float(np.abs(got[rs, cs] - r[rs, cs]).max()))


# --------------------------------------------------------------------------- vectors frozen from the reference's own source
def _ref_fixture(name):
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", name))


def _gpu_stimuli(g):
    from cardiax_b200 import stimulus
    out, i = [], 0
    while "field%d" % i in g.files:
        out.append(stimulus.Stimulus(stimulus.Protocol(*[float(x) for x in g["proto%d" % i]]), torch.as_tensor(g["field%d" % i]).cuda()))
        i += 1
    return out


@pytest.mark.parametrize("kernel", [0, 1, 3, 4])
def test_cuda_equals_reference_vectors_forward(kernel):
    """tests/golden/ref_fk_48x40.npz was produced by the UNMODIFIED reference source (cardiax/solve.py on the NumPy
    stand-in for jax, XLA's tanh): `solve.forward` with EULER and HEUN and the int-counter `_forward_euler` of
    deepx.generate.sequence.  Exact numerics reproduce it bit for bit through every kernel; fast numerics within TOL."""
    from cardiax_b200 import options, solve
    g = _ref_fixture("ref_fk_48x40.npz")
    P, stim, cps = O.Params(*g["params"]), _gpu_stimuli(g), g["checkpoints"]
    st0 = solve.State(*[torch.as_tensor(g[k + "0"]).cuda() for k in "vwu"])
    D = torch.as_tensor(g["D"]).cuda()
    options.kernel = kernel
    for name, integ in (("euler", solve.TimeIntegrator.EULER), ("heun", solve.TimeIntegrator.HEUN)):
        options.numerics = "exact"
        states = solve.forward(st0, cps, P, D, stim, float(g["dt"]), float(g["dx"]), integ)
        assert len(states) == len(cps) - 1
        for i, s in enumerate(states):
            assert_exact([x.cpu().numpy() for x in s], [g["%s_xla_%s%d" % (name, f, i + 1)] for f in "vwu"],
                         "%s checkpoint %d kernel %d" % (name, i, kernel))
        options.numerics = "fast"
        states = solve.forward(st0, cps, P, D, stim, float(g["dt"]), float(g["dx"]), integ)
        for i, s in enumerate(states):
            for f, x in zip("vwu", s):
                for tanh in ("xla", "numpy"):     # two legitimate tanh functions: fast numerics sit within TOL of both
                    assert float(np.abs(x.cpu().numpy() - g["%s_%s_%s%d" % (name, tanh, f, i + 1)]).max()) <= TOL_FAST, (name, f, i)
    options.numerics = "exact"
    s = solve._forward_euler(st0, 0, 60, P, D, stim, 0.01, 0.01)       # Python-int bounds: the int32 counter
    assert_exact([x.cpu().numpy() for x in s], [g["euler_int_xla_" + f] for f in "vwu"], "int counter")


def test_cuda_equals_reference_vectors_step_gradient_stimulate():
    """tests/golden/ref_fk_steps.npz (reference source): `solve.step` for all 16 parameter sets, `gradient`, `stimulate`."""
    from cardiax_b200 import options, solve
    g = _ref_fixture("ref_fk_steps.npz")
    stim = _gpu_stimuli(g)
    st = solve.State(*[torch.as_tensor(g[k + "0"]).cuda() for k in "vwu"])
    D = torch.as_tensor(g["D"]).cuda()
    for key, P in O.PARAMSETS.items():
        options.numerics = "exact"
        d = solve.step(st, float(g["t"]), P, D, stim, float(g["dx"]))
        assert_exact([x.cpu().numpy() for x in d], [g["d%s_xla_%s" % (f, key)] for f in "vwu"], "step paramset " + key)
    a = torch.as_tensor(g["grad_in"]).cuda()
    for axis in range(4):
        assert np.array_equal(solve.gradient(a, axis).cpu().numpy(), g["grad_axis%d" % axis])
    X = torch.as_tensor(g["stimulate_in"]).cuda()
    for t in range(24):
        assert np.array_equal(solve.stimulate(float(t), X, stim).cpu().numpy(), g["stimulate_out"][t]), t


def test_uniform_then_scar_map_of_the_same_shape_from_numpy_inputs():
    """Regression (VERDICT r1 weak 3 / ADVICE high): a homogeneous run followed, in the same process, by a scar-map run
    of the same shape -- diffusivity handed over as NumPy, so each call uploads a fresh tensor that the caching
    allocator places at the address the previous one just left -- must take the heterogeneous path."""
    from cardiax_b200 import _lib, options, solve, stimulus
    options.numerics = "exact"
    shape = (96, 640)                       # large enough for the streaming kernel, the only reader of the flag
    st, D, stim = common.random_case(shape, seed=17, n_stim=2)
    gst = [stimulus.Stimulus(stimulus.Protocol(*s.protocol), torch.as_tensor(s.field).cuda()) for s in stim]
    gstate = solve.State(*[torch.as_tensor(x).cuda() for x in st])
    Du = np.full(shape, 1e-3, np.float32)
    ref_u = C.forward_euler(st, 0, 8, P3, Du, stim, 0.01, 0.01)
    ref_s = C.forward_euler(st, 0, 8, P3, D, stim, 0.01, 0.01)
    options.kernel = 2
    for _ in range(3):
        out = solve._forward_euler(gstate, 0, 8, P3, Du, gst, 0.01, 0.01)          # NumPy in: uploaded, freed
        assert_exact([x.cpu().numpy() for x in out], ref_u, "uniform")
        del out
        out = solve._forward_euler(gstate, 0, 8, P3, D, gst, 0.01, 0.01)           # same shape, same address, a scar map
        assert _lib.last_kernel() == "fk_stream_kernel"
        assert_exact([x.cpu().numpy() for x in out], ref_s, "scar map after a uniform map")
        del out
    # the same with device tensors that are freed and re-created
    for Dn, ref in ((Du, ref_u), (D, ref_s), (Du, ref_u), (D, ref_s)):
        Dg = torch.as_tensor(Dn).cuda()
        out = solve._forward_euler(gstate, 0, 8, P3, Dg, gst, 0.01, 0.01)
        assert_exact([x.cpu().numpy() for x in out], ref, "device tensors")
        del Dg, out


def test_int_counter_and_typed_protocols_above_2_pow_24():
    """The reference's int32-counter path (deepx/generate.py:24-27, 187-196) keeps exact integer semantics where the
    float32 path rounds: counters and periods above 2^24 (VERDICT r1 missing 6; ABI v2 carries the typing)."""
    from cardiax_b200 import options, solve, stimulus
    options.numerics = "exact"
    shape = (40, 96)
    st, D, _ = common.random_case(shape, seed=31, n_stim=0)
    f = np.zeros(shape, np.float32); f[:6] = 20.0
    g = np.zeros(shape, np.float32); g[:, -9:] = -3.0
    t0 = 3 + 16777259 - 3
    protos = [O.Protocol(np.array([3], np.int32), 2, np.array([16777259], np.int32)), O.Protocol(t0 + 1, 3, 1e9)]
    stim = [O.Stimulus(p, x) for p, x in zip(protos, (f, g))]
    gst = [stimulus.Stimulus(stimulus.Protocol(*p), torch.as_tensor(x).cuda()) for p, x in zip(protos, (f, g))]
    gs = solve.State(*[torch.as_tensor(x).cuda() for x in st])
    Dg = torch.as_tensor(D).cuda()
    ref = O.forward_euler(st, t0, t0 + 8, P3, D, stim, 0.01, 0.01, counter="i32")
    for kernel in (0, 1, 2, 3, 4):
        options.kernel, options.cta_threads = kernel, (32 if kernel == 2 else 0)
        got = solve._forward_euler(gs, t0, t0 + 8, P3, Dg, gst, 0.01, 0.01)                    # Python-int bounds
        assert_exact([x.cpu().numpy() for x in got], ref, "int counter, kernel %d" % kernel)
    options.kernel, options.cta_threads = 0, 0
    X = np.zeros(shape, np.float32)
    for t in range(t0 - 2, t0 + 9):
        for tt in (int(t), float(t)):
            out = solve.stimulate(tt, torch.as_tensor(X).cuda(), gst).cpu().numpy()
            assert np.array_equal(out, O.stimulate(tt, X, stim, typed=True)), tt
    d = solve.step(gs, t0 + 3, P3, Dg, gst, 0.01)
    assert_exact([x.cpu().numpy() for x in d], O.step(st, t0 + 3, P3, D, stim, 0.01, typed=True), "step, int t")


def test_forward_dimensional_is_executed():
    """solve.forward_dimensional (cardiax/solve.py:133-165): AssertionError on a diffusivity / stimulus shape mismatch,
    checkpoints = arange(0, stop, step) in step units, init state; equals the reference-source run frozen in the oracle."""
    from cardiax_b200 import options, solve, stimulus
    options.numerics = "exact"
    D = torch.full((12, 16), 1e-3, device="cuda")
    stim = [stimulus.linear((12, 16), stimulus.Direction.NORTH, 0.3, 20.0, stimulus.Protocol(0, 2, 1e9))]
    with pytest.raises(AssertionError):
        solve.forward_dimensional((0.12, 0.17), 0.4, 0.1, P3, D, stim, 0.01, 0.01)
    with pytest.raises(AssertionError):
        solve.forward_dimensional((0.121, 0.161), 0.4, 0.1, P3, D, [stimulus.linear((12, 20), 0, 0.3, 20.0, stimulus.Protocol(0, 2, 1e9))], 0.01, 0.01)
    got = solve.forward_dimensional((0.121, 0.161), 0.401, 0.101, P3, D, stim, 0.01, 0.01)
    cps = np.arange(0, int(0.401 / 0.01), int(0.101 / 0.01))
    assert len(got) == len(cps) - 1
    ost = [O.Stimulus(O.Protocol(0, 2, 1e9), stim[0].field.cpu().numpy())]
    s = O.init((12, 16))
    for i in range(len(cps) - 1):
        s = O.forward_euler(s, float(cps[i]), float(cps[i + 1]), P3, np.full((12, 16), 1e-3, np.float32), ost, 0.01, 0.01)
        assert_exact([x.cpu().numpy() for x in got[i]], s, "forward_dimensional checkpoint %d" % i)


@pytest.mark.parametrize("uniform", [False, True])
@pytest.mark.parametrize("pset", sorted(O.PARAMSETS))
def test_fast_all_paramsets_on_the_streaming_kernel(pset, uniform):
    """Fast numerics against the oracle for ALL 16 parameter sets -- incl. sets 2 and 7, whose k = 1e5 / 1e6 and
    V_csi = 1e6 drive ex2.approx to +inf and rcp.approx(inf) to 0 (j_si == 0, as in the reference where tanh == -1) -- on
    the streaming kernel's heterogeneous-D and uniform-D instantiations."""
    from cardiax_b200 import _lib
    shape = (96, 640)
    st, D = common.smooth_case(shape, seed=12)
    if uniform:
        D = np.full(shape, 1e-3, np.float32)
    _, _, stim = common.random_case(shape, seed=12, n_stim=2)
    P = O.PARAMSETS[pset]
    ref = C.forward_euler(st, 0, 60, P, D, stim, 0.01, 0.01)
    ref64 = C.forward_euler(st, 0, 60, P, D, stim, 0.01, 0.01, dtype=np.float64)
    got = run_gpu(st, 0, 60, P, D, stim, numerics="fast", kernel=2, steps_per_launch=2)
    assert _lib.last_kernel() == "fk_stream_kernel"
    for name, g, a, b in zip("vwu", got, ref, ref64):
        assert np.isfinite(g).all(), name
        assert np.abs(g - a).max() <= TOL_FAST, (pset, name, float(np.abs(g - a).max()))
        # the float32 oracle itself sits 9e-7 ... 1e-5 from the float64 twin over the 16 sets of this case (its gap is a
        # draw, not a bound: tools measured on the CPU emulation), so the envelope has that floor
        assert np.abs(g - b).max() <= max(2 * np.abs(a - b).max() + 2e-6, 1e-5), (pset, name)


def test_config2_real_protocol_across_the_second_stimulus():
    """BASELINE config 2 as the reference runs it (experiments/generate_fd_data_256.py:11-12, 22-40): 512 x 512 scar map,
    S2 at step 40 000, segments of 500 steps -- up to step 40 500, ACROSS S2.  Exact numerics: SHA-256 of v, w, u equal the
    C oracle's at steps 39 500, 40 000, 40 002 and 40 500.  Fast numerics: as close to the float64 run as the float32
    oracle is (the fixture holds both), and S2 has visibly fired.  Fixture + generator: tests/golden/make_config2.py."""
    import hashlib
    import os
    from cardiax_b200 import _lib, options, solve, stimulus
    from tests.golden import make_config2 as M
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "fk_512_config2_s2.npz"))
    D, stim = M.inputs()
    assert (M.sha(D), M.sha(stim[0].field), M.sha(stim[1].field)) == (str(g["D_sha"]), str(g["s1_sha"]), str(g["s2_sha"])), \
        "the regenerated inputs differ from the ones the fixture was made from (scipy / numpy version?)"
    gst = [stimulus.Stimulus(stimulus.Protocol(*s.protocol), torch.as_tensor(s.field).cuda()) for s in stim]
    Dg = torch.as_tensor(D).cuda()
    marks = [int(m) for m in g["marks"]]
    for numerics in ("exact", "fast"):
        options.numerics = numerics
        s, t = solve.init(M.SHAPE), 0
        for m in marks:
            while t < m:                                   # the reference's 500-step segments, float counters
                nxt = min(m, t + 500)
                s = solve._forward_euler(s, float(t), float(nxt), P3, Dg, gst, 0.01, 0.01)
                t = nxt
            host = [x.cpu().numpy() for x in s]
            if numerics == "exact":
                got = [hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest() for x in host]
                assert got == [str(x) for x in g["sha_%d" % m]], "step %d" % m
            elif m >= 40002:
                for name, x in zip("vwu", host):
                    f32, f64 = g["%s_f32_%d" % (name, m)], g["%s_f64_%d" % (name, m)]
                    sub = x[::4, ::4]
                    gap, ours = np.abs(f32 - f64), np.abs(sub - f64)
                    print("config 2 step %d %s: |f32 - f64| max %.3g p99 %.3g, |fast - f64| max %.3g p99 %.3g" % (
                        m, name, gap.max(), np.percentile(gap, 99), ours.max(), np.percentile(ours, 99)))
                    assert np.percentile(ours, 99) <= 3 * np.percentile(gap, 99) + 1e-5, (m, name)
                    assert ours.max() <= 3 * gap.max() + 1e-4, (m, name)
        assert _lib.last_kernel() == "fk_resident_kernel"
    # S2 did fire: the state just after it differs from the state just before by a stimulus-sized jump
    assert float(np.abs(g["u_f32_40002"] - g["u_f32_40500"]).max()) > 0.05
