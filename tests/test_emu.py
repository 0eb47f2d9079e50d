"""CPU emulation of the CUDA kernel bodies (tests/emu) against the oracle: index, boundary, ring and launch logic of
the product's kernels, checked without a GPU.  EXACT mode must be bit-identical; fast mode within tolerance."""
import os

import numpy as np
import pytest

import oracle as O
from oracle import c_oracle as C
from tests import common
from tests.emu import emu

P3 = O.PARAMSETS["3"]


def _exact(shape, T, nsteps=7, n_stim=3, seed=0, **kw):
    st, D, stim = common.random_case(shape, seed=seed, n_stim=n_stim)
    ref = C.forward_euler(st, 0, nsteps, P3, D, stim, 0.01, 0.01)
    got, info = emu.euler(st, 0, nsteps, P3, D, stim, 0.01, 0.01, exact=True, T=T, **kw)
    for name, a, b in zip("vwu", got, ref):
        assert not np.isnan(a).any(), name
        assert np.array_equal(a, b), (name, float(np.abs(a - b).max()))
    return info


@pytest.mark.parametrize("shape,T", [((37, 53), 1), ((37, 53), 2), ((37, 53), 3), ((5, 7), 2), ((3, 3), 1), ((4, 9), 4),
                                     ((70, 35), 2), ((33, 130), 8)])
def test_tile_kernel_exact(shape, T):
    info = _exact(shape, T, kernel=1)
    assert info[1] == 0 and info[0] > 0


@pytest.mark.parametrize("shape,T,nt,rh", [((72, 160), 1, 32, 16), ((72, 160), 2, 32, 16), ((96, 288), 3, 32, 20),
                                           ((80, 300), 2, 64, 0), ((130, 516), 4, 64, 17), ((56, 1100), 2, 0, 0),
                                           ((200, 64), 2, 0, 0)])
def test_stream_plus_frame_exact(shape, T, nt, rh):
    info = _exact(shape, T, nsteps=2 * T + 1, kernel=2, cta_threads=nt, rows_per_cta=rh)
    assert info[1] > 0  # the streaming kernel really ran; the odd tail step goes through the tile kernel


@pytest.mark.parametrize("shape", [(12, 16), (3, 4), (5, 8), (40, 64), (9, 132), (130, 20)])
def test_wide_kernel_exact(shape):
    info = _exact(shape, 1, nsteps=6, kernel=3)
    assert info == (0, 6000)   # six launches of the wide kernel, nothing else


@pytest.mark.parametrize("nc", [0, 1, 2, 4])
@pytest.mark.parametrize("shape,nsteps,tiles", [
    ((12, 16), 6, (0, 0)), ((3, 4), 5, (0, 0)), ((64, 96), 7, (0, 0)), ((64, 96), 7, (2, 3)), ((64, 96), 7, (4, 1)),
    ((40, 64), 9, (3, 2)), ((37, 52), 8, (2, 2)), ((130, 20), 6, (0, 0)), ((9, 132), 6, (0, 0)), ((100, 200), 6, (5, 5)),
    ((64, 96), 1, (2, 3)), ((64, 96), 2, (1, 1)), ((128, 128), 5, (0, 0)), ((33, 72), 6, (4, 9)), ((7, 8), 5, (1, 1))])
def test_resident_kernel_exact(shape, nsteps, tiles, nc):
    """Whole call in one resident launch (tiles in shared memory, halos through tagged mailbox records): any tile grid,
    uneven tiles, ring-only tiles, 1/2/4 cells per thread, all four physical edges (edge cells one per thread), stimuli
    -- bit-identical to the oracle."""
    info = _exact(shape, 1, nsteps=nsteps, kernel=4, tiles=tiles, nc=nc)
    assert info == (0, 1000000)   # one launch, nothing else


@pytest.mark.parametrize("nc", [0, 1, 2, 4])
@pytest.mark.parametrize("shape,nsteps,tiles", [
    ((12, 16), 6, (0, 0)), ((64, 96), 7, (0, 0)), ((64, 96), 7, (2, 3)), ((64, 96), 7, (4, 1)), ((64, 96), 7, (1, 5)),
    ((40, 64), 9, (3, 2)), ((37, 52), 8, (2, 2)), ((130, 20), 6, (0, 0)), ((128, 128), 5, (0, 0)), ((128, 128), 5, (4, 4)),
    ((96, 160), 6, (3, 5)), ((7, 8), 5, (1, 1)), ((64, 96), 1, (2, 3))])
def test_cluster_transport_exact(shape, nsteps, tiles, nc):
    """The resident kernel with the tissue as ONE thread-block cluster: ring cells stored straight into the neighbours'
    halos (distributed shared memory on the device, the neighbour's buffer here), no mailboxes -- bit-identical to the
    oracle for any grid of <= 16 tiles, uneven tiles, every cells-per-thread, all physical edges, stimuli."""
    info = _exact(shape, 1, nsteps=nsteps, kernel=5, tiles=tiles, nc=nc)
    assert info == (0, 1000000)


def test_cluster_transport_batch_and_limits():
    shape, batch = (40, 48), 3
    cases = [common.random_case(shape, seed=20 + b, n_stim=2) for b in range(batch)]
    st = [np.stack([c[0][k] for c in cases]) for k in range(3)]
    D = np.stack([c[1] for c in cases])
    stims = [c[2] for c in cases]
    got, info = emu.euler(st, 0, 6, P3, D, stims, 0.01, 0.01, exact=True, kernel=5, tiles=(2, 2))
    assert info == (0, 1000000)
    for b in range(batch):
        ref = C.forward_euler(cases[b][0], 0, 6, P3, cases[b][1], cases[b][2], 0.01, 0.01)
        for a, r in zip(got, ref):
            assert np.array_equal(a[b], r)
    # more than 16 tiles is not a cluster
    with pytest.raises(Exception):
        emu.euler([x[0] for x in st], 0, 6, P3, D[0], stims[0], 0.01, 0.01, exact=True, kernel=5, tiles=(5, 4))


@pytest.mark.parametrize("shape,tiles,edge,nc", [((96, 160), (5, 6), (12, 4), 4), ((96, 160), (5, 6), (12, 4), 1),
                                                 ((120, 96), (4, 3), (20, 6), 2), ((120, 96), (6, 1), (10, 0), 4)])
def test_resident_kernel_uneven_edge_tiles(shape, tiles, edge, nc):
    """Smaller tiles at the tissue's edges (their one-sided formulas cost more): same bits."""
    info = _exact(shape, 1, nsteps=6, kernel=4, tiles=tiles, nc=nc, edge_tile=edge)
    assert info == (0, 1000000)


@pytest.mark.parametrize("shape,tiles,nc", [((64, 96), (2, 3), 4), ((40, 64), (3, 2), 2), ((33, 72), (1, 1), 1)])
def test_resident_kernel_maps_in_global_memory(shape, tiles, nc):
    """Larger tissues keep only u, v, w in shared memory and read D, D_x, D_y from L2: same bits."""
    info = _exact(shape, 1, nsteps=6, kernel=4, tiles=tiles, nc=nc, maps_global=1)
    assert info == (0, 1000000)


def test_resident_kernel_exact_batched_per_tissue_inputs():
    shape, batch = (40, 48), 3
    cases = [common.random_case(shape, seed=20 + b, n_stim=2) for b in range(batch)]
    st = [np.stack([c[0][k] for c in cases]) for k in range(3)]
    D = np.stack([c[1] for c in cases])
    stims = [c[2] for c in cases]
    got, info = emu.euler(st, 0, 6, P3, D, stims, 0.01, 0.01, exact=True, kernel=4, tiles=(2, 2))
    assert info == (0, 1000000)
    for b in range(batch):
        ref = C.forward_euler(cases[b][0], 0, 6, P3, cases[b][1], cases[b][2], 0.01, 0.01)
        for a, r in zip(got, ref):
            assert np.array_equal(a[b], r)


def test_resident_planner():
    p = emu.plan_resident(512, 512)
    assert p and p["ntr"] * p["ntc"] <= 148 and p["smem_bytes"] <= 227 * 1024 and p["threads"] % 32 == 0
    assert p["nc"] == 4 and p["two_pass"] == 0            # (the two-pass step is a compile-time experiment, off)
    p = emu.plan_resident(1200, 1200)                      # the reference's data-generation tissue: 40 MB of state + maps
    assert p and p["maps_in_l2"] == 1 and p["smem_bytes"] <= 227 * 1024   # do not fit 148 x 227 KB; u, v, w alone do
    assert emu.plan_resident(1600, 1600) is None
    assert emu.plan_resident(64, 66) is None              # rows must be whole groups of 4 cells
    p = emu.plan_resident(256, 256, batch=8)
    assert p and p["ntr"] * p["ntc"] * 8 <= 148


def test_small_tissue_defaults_and_fast_matches_other_kernels():
    (st, D) = common.smooth_case((64, 96), seed=6)
    _, _, stim = common.random_case((64, 96), seed=6, n_stim=2)
    a, info = emu.euler(st, 0, 7, P3, D, stim, 0.01, 0.01, exact=False)
    assert info == (0, 1000000)   # >= 4 steps of a small tissue: the resident kernel
    a3, info = emu.euler(st, 0, 3, P3, D, stim, 0.01, 0.01, exact=False)
    assert info == (0, 3000)      # fewer: one wide launch per step
    w7, info = emu.euler(st, 0, 7, P3, D, stim, 0.01, 0.01, exact=False, kernel=3)
    assert info == (0, 7000)
    for x, y in zip(a, w7):
        assert np.array_equal(x, y)
    b, _ = emu.euler(st, 0, 7, P3, D, stim, 0.01, 0.01, exact=False, kernel=1, T=2)
    c, _ = emu.euler(st, 0, 7, P3, D, stim, 0.01, 0.01, exact=False, kernel=2, T=2, cta_threads=32)
    for x, y, z in zip(a, b, c):
        assert np.array_equal(x, y) and np.array_equal(x, z)


def test_stream_thread_order_independent():
    """Threads of a block run in reverse order inside every iteration: any same-iteration hazard on the rings shows."""
    _exact((72, 160), 2, kernel=2, cta_threads=32, rows_per_cta=16, reverse=1)
    _exact((96, 288), 3, kernel=2, cta_threads=64, rows_per_cta=20, reverse=1)


@pytest.mark.parametrize("shape,T,nt,rh,reverse", [((96, 400), 2, 32, 40, 0), ((96, 400), 2, 32, 40, 1),
                                                   ((70, 264), 1, 32, 27, 0), ((120, 520), 3, 64, 50, 1),
                                                   ((150, 64), 2, 0, 0, 0)])
def test_stream_steady_state_body_exact(shape, T, nt, rh, reverse):
    """No stimulus: after the pipeline fill the kernel runs its unrolled steady-state body (static ring slots, rotating
    u_x window, swapped ring halves), several bodies per CTA, in edge strips and interior strips."""
    info = _exact(shape, T, nsteps=2 * T, n_stim=0, kernel=2, cta_threads=nt, rows_per_cta=rh, reverse=reverse)
    assert info[1] == 2


@pytest.mark.parametrize("shape,T,nt,rh,n_stim,uniform", [
    ((8, 64), 1, 0, 0, 0, 0), ((9, 64), 1, 0, 0, 2, 0), ((16, 48), 2, 32, 0, 0, 1), ((24, 64), 3, 32, 24, 0, 0),
    ((33, 64), 2, 32, 8, 0, 0), ((41, 64), 2, 32, 16, 2, 0), ((40, 64), 1, 32, 40, 0, 1), ((72, 160), 2, 32, 16, 0, 1),
    ((44, 96), 1, 32, 40, 0, 0), ((130, 516), 4, 64, 17, 0, 0), ((57, 132), 2, 32, 20, 3, 0), ((200, 64), 2, 0, 0, 0, 1)])
def test_stream_owns_the_physical_top_and_bottom_edges(shape, T, nt, rh, n_stim, uniform):
    """No frame tiles any more: the first / last row chunk of the streaming kernel evaluates the reference's one-sided
    formulas at the tissue's top / bottom edge (first row deferred by one iteration, last rows from saved u_x), for
    single-chunk tissues, short last chunks, every T, with and without stimuli -- bit for bit, zero tile launches."""
    st, D, stim = common.random_case(shape, seed=7, n_stim=n_stim)
    if uniform:
        D = np.full(shape, 1e-3, np.float32)
    n = 2 * T
    ref = C.forward_euler(st, 0, n, P3, D, stim, 0.01, 0.01)
    got, info = emu.euler(st, 0, n, P3, D, stim, 0.01, 0.01, exact=True, T=T, kernel=2, cta_threads=nt, rows_per_cta=rh,
                          uniform=uniform)
    assert info == (0, 2)
    for name, a, b in zip("vwu", got, ref):
        assert np.array_equal(a, b), (name, np.argwhere(a != b)[:4].tolist())


@pytest.mark.parametrize("lo,hi,pt,pb,row0,row1", [(0, 48, 1, 0, 0, 44), (0, 48, 1, 0, 32, 40), (32, 80, 0, 1, 4, 48),
                                                   (32, 80, 0, 1, 8, 16), (32, 80, 0, 1, 16, 48), (16, 64, 0, 0, 4, 44)])
def test_stream_row_windows_of_a_slab(lo, hi, pt, pb, row0, row1):
    """fk_euler_rows building block: a window of output rows of a slab buffer, physical edge on one side or none."""
    H, W = 80, 160
    st, D, stim = common.random_case((H, W), seed=12, n_stim=2)
    ref = C.forward_euler(st, 0, 1, P3, D, stim, 0.01, 0.01)
    sub = [x[lo:hi] for x in st]
    ss = [O.Stimulus(s.protocol, s.field[lo:hi]) for s in stim]
    got, info = emu.euler(sub, 0, 1, P3, D[lo:hi], ss, 0.01, 0.01, exact=True, T=1, kernel=2, cta_threads=32, phys_top=pt,
                          phys_bottom=pb, row0=row0, row1=row1)
    assert info == (0, 1)
    for g, r in zip(got, ref):
        assert np.array_equal(g[row0:row1], r[lo + row0:lo + row1])


def test_stream_steady_state_uniform_fast_matches_tiles():
    st, _ = common.smooth_case((96, 400), seed=3)
    D = np.full((96, 400), 1e-3, np.float32)
    a, _ = emu.euler(st, 0, 4, P3, D, [], 0.01, 0.01, exact=False, T=2, kernel=2, cta_threads=32, rows_per_cta=40, uniform=1)
    b, _ = emu.euler(st, 0, 4, P3, D, [], 0.01, 0.01, exact=False, T=2, kernel=1)
    e, _ = emu.euler(st, 0, 4, P3, D, [], 0.01, 0.01, exact=True, T=2, kernel=2, cta_threads=32, rows_per_cta=40, uniform=1)
    ref = C.forward_euler(st, 0, 4, P3, D, [], 0.01, 0.01)
    for x, y, z, r in zip(a, b, e, ref):
        assert np.array_equal(x, y) and np.array_equal(z, r)


def test_stream_uniform_diffusivity_path():
    st, _, stim = common.random_case((72, 288), seed=4)
    D = np.full((72, 288), 1e-3, np.float32)
    ref = C.forward_euler(st, 0, 6, P3, D, stim, 0.01, 0.01)
    got, info = emu.euler(st, 0, 6, P3, D, stim, 0.01, 0.01, exact=True, T=2, kernel=2, cta_threads=32, uniform=1)
    assert info[1] == 3
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("pset", sorted(O.PARAMSETS))
def test_all_paramsets_exact(pset):
    st, D, stim = common.random_case((48, 96), seed=5)
    ref = C.forward_euler(st, 0, 5, O.PARAMSETS[pset], D, stim, 0.01, 0.01)
    got, _ = emu.euler(st, 0, 5, O.PARAMSETS[pset], D, stim, 0.01, 0.01, exact=True, T=2, kernel=2, cta_threads=32)
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)


def test_fast_mode_tolerance_and_tiling_independence():
    (st, D) = common.smooth_case((136, 520), seed=4)
    _, _, stim = common.random_case((136, 520), seed=4, n_stim=2)
    ref = C.forward_euler(st, 0, 12, P3, D, stim, 0.01, 0.01)
    a, _ = emu.euler(st, 0, 12, P3, D, stim, 0.01, 0.01, exact=False, kernel=1, T=1)
    for x, y in zip(a, ref):
        assert float(np.abs(x - y).max()) <= 2e-5
    for kw in (dict(kernel=1, T=3), dict(kernel=2, T=2, cta_threads=64, rows_per_cta=24), dict(kernel=2, T=4, cta_threads=128)):
        b, _ = emu.euler(st, 0, 12, P3, D, stim, 0.01, 0.01, exact=False, **kw)
        for x, y in zip(a, b):
            assert np.array_equal(x, y), kw


def test_rhs_mode_is_solve_step():
    st, D, stim = common.random_case((24, 40), seed=6, n_stim=3)
    for t in (0, 1, 3, 4):
        got, _ = emu.euler(st, t, t + 1, P3, D, stim, 0.0, 0.01, exact=True, rhs=True)
        ref = O.step(st, t, P3, D, stim, 0.01)
        for a, b in zip(got, ref):
            assert np.array_equal(a, b)


def test_batch_and_per_tissue_stimuli():
    shape, B = (40, 160), 3
    cases = [common.random_case(shape, seed=30 + b) for b in range(B)]
    per = [[O.Stimulus(O.Protocol(b, 2, 4 + b), s.field) for s in cases[b][2]] for b in range(B)]
    st = [np.stack([c[0][k] for c in cases]) for k in range(3)]
    D = np.stack([c[1] for c in cases])
    got, _ = emu.euler(st, 0, 9, P3, D, per, 0.01, 0.01, exact=True, T=2, kernel=2, cta_threads=32)
    for b in range(B):
        ref = C.forward_euler(cases[b][0], 0, 9, P3, cases[b][1], per[b], 0.01, 0.01)
        for a, r in zip(got, ref):
            assert np.array_equal(a[b], r)


def test_slab_rows_with_halo_match_whole_tissue():
    """Row-slab decomposition: a slab buffer with 4T halo rows and non-physical edges reproduces the whole tissue's rows."""
    shape, T = (96, 160), 2
    st, D, stim = common.random_case(shape, seed=8)
    ref = C.forward_euler(st, 0, T, P3, D, stim, 0.01, 0.01)
    F = 4 * T
    for (r0, r1, top, bot) in ((0, 40, 1, 0), (40, 70, 0, 0), (70, 96, 0, 1)):
        a, b = max(0, r0 - F), min(96, r1 + F)
        sub = [x[a:b] for x in st]
        sstim = [O.Stimulus(s.protocol, s.field[a:b]) for s in stim]
        for kernel in (1, 2):
            got, _ = emu.euler(sub, 0, T, P3, D[a:b], sstim, 0.01, 0.01, exact=True, T=T, kernel=kernel, cta_threads=32,
                               phys_top=top, phys_bottom=bot)
            for g, r in zip(got, ref):
                assert np.array_equal(g[r0 - a:r1 - a], r[r0:r1]), (r0, r1, kernel)


def test_row_windows_compose_to_the_whole_slab():
    """fk_euler_rows building block: edge bands + interior, written into one output, equal a whole-tissue launch."""
    shape, T = (120, 160), 2
    st, D, stim = common.random_case(shape, seed=9)
    ref = C.forward_euler(st, 0, T, P3, D, stim, 0.01, 0.01)
    F = 4 * T
    # middle rank of a 3-way split with a deeper (2F) halo: buffer rows 24..96, non-physical edges, owns 40..80
    a, b = 24, 96
    sub = [x[a:b] for x in st]
    sstim = [O.Stimulus(s.protocol, s.field[a:b]) for s in stim]
    out = [np.full((b - a,) + shape[1:], np.nan, np.float32) for _ in range(3)]
    for (r0, r1) in ((16, 16 + F), (56 - F, 56), (16 + F, 56 - F)):   # top band, bottom band, interior (buffer coords)
        got, _ = emu.euler(sub, 0, T, P3, D[a:b], sstim, 0.01, 0.01, exact=True, T=T, kernel=2, cta_threads=32,
                           phys_top=0, phys_bottom=0, row0=r0, row1=r1)
        for o, g in zip(out, got):
            assert np.isnan(g[:r0]).all() and np.isnan(g[r1:]).all()      # nothing outside the window is written
            o[r0:r1] = g[r0:r1]
    for o, r in zip(out, ref):
        assert np.array_equal(o[16:56], r[40:80])
    # top rank: physical top edge, window starting at row 0
    got, _ = emu.euler([x[:60] for x in st], 0, T, P3, D[:60], [O.Stimulus(s.protocol, s.field[:60]) for s in stim], 0.01,
                       0.01, exact=True, T=T, kernel=2, cta_threads=32, phys_top=1, phys_bottom=0, row0=0, row1=44)
    for g, r in zip(got, ref):
        assert np.array_equal(g[:44], r[:44])


def test_device_schedule_function_equals_oracle():
    rng = np.random.default_rng(1)
    for _ in range(200):
        start, dur, per = int(rng.integers(0, 50)), int(rng.integers(1, 6)), float(rng.choice([7, 50, 400, 1e6, 1e9]))
        for t in range(0, 100):
            assert emu.stim_active(t, start, dur, per) == O.stimulus_active(t, O.Protocol(start, dur, per))


@pytest.mark.parametrize("shape,n", [((40, 72), 3), ((33, 65), 2), ((97, 130), 2), ((9, 12), 4)])
def test_fused_heun_exact(shape, n):
    """fk_driver.h: drive_heun -- ONE tile launch per Heun step (predictor and corrector are the launch's two levels),
    bit-identical to the oracle's literal step_heun loop (cardiax/solve.py:73-85, 103-111); several tiles per tissue,
    remainder tiles, stimuli switching on and off between steps (both stages of a step see the same counter)."""
    st, D, stim = common.random_case(shape, seed=5, n_stim=3)
    ref = O.forward_heun(st, 2, 2 + n, O.PARAMSETS["3"], D, stim, 0.01, 0.01)
    got, launches = emu.heun(st, 2, 2 + n, O.PARAMSETS["3"], D, stim, 0.01, 0.01)
    assert launches == n
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("shape,force_stream", [((64, 96), False), ((80, 128), True), ((37, 50), False)])
def test_fast_heun_through_the_euler_kernels(shape, force_stream):
    """fk_driver.h: drive_heun_fast -- y + (E(E(y)) - y) / 2 with E the wide / streaming Euler kernels: one two-step call
    when no stimulus is active at t or t + 1, two single-step calls at the same counter otherwise; the closing pass folded
    into the last launch's store gives the same bits as the separate combine pass; within the fast tolerance of the
    oracle's literal Heun loop (cardiax/solve.py:73-85).  (37, 50): W % 4 != 0 -> general tiles + combine pass.)"""
    st, D = common.smooth_case(shape, seed=4)
    stim = [O.linear(shape, 0, 0.3, 20.0, O.Protocol(2, 2, 50))]
    n = 7
    ref = O.forward_heun(st, 0, n, O.PARAMSETS["3"], D, stim, 0.01, 0.01)
    a, ia = emu.heun_fast(st, 0, n, O.PARAMSETS["3"], D, stim, 0.01, 0.01, fold=True, force_stream=force_stream)
    b, ib = emu.heun_fast(st, 0, n, O.PARAMSETS["3"], D, stim, 0.01, 0.01, fold=False, force_stream=force_stream)
    for x, y, r in zip(a, b, ref):
        assert np.array_equal(x, y)
        assert np.abs(x - r).max() <= 2e-5 * max(1.0, np.abs(r).max())
    assert np.abs(a[2] - st.u).max() > 1e-3
    assert ib["combine"] == n
    if shape[1] % 4 == 0:
        assert ia["combine"] == 0 and ia["tile"] == 0
        # quiet steps t = 0, 4, 5, 6 -> one two-step call; t = 1, 2, 3 (stimulus active at t or t + 1) -> two one-step calls
        assert (ia["stream"] if force_stream else ia["wide"]) == (4 * 1 + 3 * 2 if force_stream else 4 * 2 + 3 * 2)
    else:
        assert ia["combine"] == n and ia["tile"] > 0


@pytest.mark.parametrize("exact", [True, False])
def test_stream_kernel_mirrors_the_slab_edge_rows(exact):
    """Row-slab halo exchange fused into the step kernel (FK_STORE_MIRROR): a row-window launch also stores its first /
    last `band` result rows into the neighbours' arrays at the requested rows -- and nothing else there."""
    H, W, T, band = 120, 160, 2, 16
    st, D, stim = common.random_case((H, W), seed=21, n_stim=2)
    row0, row1 = 8, 112                       # a middle slab: halos of 4T rows on both sides are inputs only
    up = [np.full((90, W), -7.0, np.float32) for _ in range(3)]
    down = [np.full((70, W), -9.0, np.float32) for _ in range(3)]
    emu.set_mirror(up, down, (row0, row1 - band), (row0 + band, row1), (90 - band, 3))
    got, _ = emu.euler(st, 5, 5 + T, P3, D, stim, 0.01, 0.01, exact=exact, T=T, kernel=2, cta_threads=32, phys_top=0,
                       phys_bottom=0, row0=row0, row1=row1)
    assert emu.mirror_was_fused()
    ref, _ = emu.euler(st, 5, 5 + T, P3, D, stim, 0.01, 0.01, exact=exact, T=T, kernel=2, cta_threads=32, phys_top=0,
                       phys_bottom=0, row0=row0, row1=row1)
    assert not emu.mirror_was_fused()
    for k in range(3):
        assert np.array_equal(got[k][row0:row1], ref[k][row0:row1])
        assert np.array_equal(up[k][90 - band:], ref[k][row0:row0 + band]) and np.all(up[k][:90 - band] == -7.0)
        assert np.array_equal(down[k][3:3 + band], ref[k][row1 - band:row1])
        assert np.all(down[k][:3] == -9.0) and np.all(down[k][3 + band:] == -9.0)


def _typed_cases():
    protos = [O.Protocol(np.array([3], np.int32), 2, np.array([16777259], np.int32)), O.Protocol(0, 2, 33554467),
              O.Protocol(5, 4, np.array([123456789], np.int32)), O.Protocol(7, 2, 1e9), O.Protocol(7.0, 2, 400), O.Protocol(0, 2, 50),
              O.Protocol(3.0, 2.0, 16777259.0), O.Protocol(2, 3.0, 9), O.Protocol(np.array([1], np.int32), 2.0, 6)]
    for proto in protos:
        s0, p0 = int(np.asarray(proto.start).reshape(-1)[0]), int(float(np.asarray(proto.period).reshape(-1)[0]))
        ts = sorted(set(list(range(0, 14)) + [s0 + k * p0 + d for k in (1, 2) for d in range(-5, 6)]))
        for t in ts:
            if 0 <= t < 2 ** 31 - 8:
                yield proto, t


def test_typed_stimulus_schedule_equals_the_oracle_above_2_pow_24():
    """fk::stim_active_typed (every kernel's schedule predicate, ABI v2) == the oracle's typed restatement -- which
    tests/test_reference_pin.py holds to the reference's own `stimulate` -- for int32 and float32 counters, int / float /
    shape-(1,) int32 protocol entries, periods and counters above 2^24 where the two typings part."""
    n = 0
    for proto, t in _typed_cases():
        for tt in (int(t), np.int32(t), float(t), np.float32(t)):
            assert emu.stim_active_typed(tt, proto) == O.stimulus_active_typed(tt, proto), (proto, tt)
            n += 1
    assert n > 1000
    # the all-float typing is the fp32 predicate the kernels used before ABI v2
    for proto, t in _typed_cases():
        fl = [float(np.asarray(x).reshape(-1)[0]) for x in proto]
        assert emu.stim_active_typed(float(t), O.Protocol(*fl)) == emu.stim_active(float(t), *fl)


@pytest.mark.parametrize("kernel", [1, 2, 3, 4])
def test_int_counter_run_matches_the_typed_oracle(kernel):
    """`deepx.generate.sequence`'s typing end to end: int bounds, int32-array protocols, a period above 2^24 whose
    second pulse the float32 typing would misplace; every kernel, exact numerics, bit for bit."""
    shape = (24, 64) if kernel != 2 else (40, 96)
    st, D, _ = common.random_case(shape, seed=31, n_stim=0)
    f = np.zeros(shape, np.float32); f[:6] = 20.0
    g = np.zeros(shape, np.float32); g[:, -9:] = -3.0
    t0 = 3 + 16777259 - 3
    stim = [O.Stimulus(O.Protocol(np.array([3], np.int32), 2, np.array([16777259], np.int32)), f),
            O.Stimulus(O.Protocol(t0 + 1, 3, 1e9), g)]
    ref = O.forward_euler(st, t0, t0 + 8, P3, D, stim, 0.01, 0.01, counter="i32")
    got, _ = emu.euler(st, t0, t0 + 8, P3, D, stim, 0.01, 0.01, exact=True, kernel=kernel, T=2 if kernel in (1, 2) else 0,
                       cta_threads=32 if kernel == 2 else 0, typed=True)
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)
    # (the reference's float32 counter cannot even run here: above 2^24 `i + 1 == i` and its fori_loop never ends)
    with pytest.raises(OverflowError):
        O.forward_euler(st, float(t0), float(t0 + 8), P3, D, stim, 0.01, 0.01, counter="f32")
    assert any(float(np.abs(a - b).max()) > 1e-3 for a, b in zip(ref, O.forward_euler(st, t0, t0 + 8, P3, D, [], 0.01, 0.01, counter="i32")))


def test_quiet_launches_are_never_wrongly_declared():
    """fk_core.h: stims_quiet -- a launch is run without the stimulus machinery only if NO step of it can fire: compared
    with the typed schedule evaluated step by step (float32 and int32 counters, small and huge periods, launches before,
    across and after the pulses, counters beyond 2^24 where float32 counters stall)."""
    rng = np.random.default_rng(5)
    protos = [(0, 2, 1e9), (40000, 2, 1e9), (3, 5, 20), (1, 2, 400), (25001, 2, 123456789), (7, 3, 50.0), (0.0, 2.0, 1e9),
              (np.int32(1), 2, np.int32(999999999)), (16777210, 2, 1e9), (16777300, 2, 40), (5, 2, 1)]
    said_quiet = 0
    for proto in protos:
        starts = [0, 1, 2, 5, 19, 38, 399, 400, 39990, 40001, 16777200, 16777290, int(rng.integers(0, 100000))]
        for t0 in starts:
            for nsteps in (1, 2, 4, 16, 500):
                for typed_int in (False, True):
                    ts = [(t0 + k) if typed_int else float(t0 + k) for k in range(nsteps)]
                    active = any(emu.stim_active_typed(t if typed_int else np.float32(t), proto) for t in ts)
                    quiet = emu.stims_quiet(t0, nsteps, proto)
                    assert not (quiet and active), (proto, t0, nsteps, typed_int)
                    said_quiet += quiet
    assert said_quiet > 300     # ... and it does recognise the sleeping launches


def test_balanced_row_chunks():
    """The chunks at the physical top / bottom edge get fewer rows than the others (their special iterations run in the
    general body), the number of chunks stays, every row is covered once; a run on such a plan is bit-exact."""
    base = emu.plan_rows(4096, 4096, T=2)
    bal = emu.plan_rows(4096, 4096, T=2, top_off=16, bot_off=28)
    assert base and bal and len(base) == len(bal) and bal[0] == 0 and bal[-1] == 4096
    rows = [b - a for a, b in zip(bal, bal[1:])]
    inner = rows[1:-1]
    assert max(inner) - min(inner) <= 4 and all(r % 4 == 0 for r in rows[:-1])
    assert rows[0] == max(inner) - 16 and max(inner) - 28 - 4 < rows[-1] <= max(inner) - 28
    assert sorted(inner, reverse=True) == inner                      # the taller chunks first
    assert emu.plan_rows(256, 256, batch=128, T=2, top_off=16, bot_off=28) is not None
    for sh, T in (((200, 64), 2), ((333, 96), 3), ((200, 64), 1)):   # small chunks: offsets of 4 rows exercise the same code
        os.environ["FK_TOP_OFF"], os.environ["FK_BOT_OFF"] = "4", "4"
        try:
            info = _exact(sh, T, nsteps=2 * T, kernel=2)
        finally:
            del os.environ["FK_TOP_OFF"], os.environ["FK_BOT_OFF"]
        assert info[1] > 0
