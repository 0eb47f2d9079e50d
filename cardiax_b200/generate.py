"""The data-generation driver of the reference (deepx/generate.py) on the B200 solver.

Same function names, arguments and defaults as ``deepx.generate`` -- ``random_protocol``, ``random_*_stimulus``,
``random_stimulus``, ``random_diffusivity``, ``random_sequence``, ``sequence`` -- with two differences that SURVEY.md
section 8f-2 allows: ``rng`` is a NumPy ``Generator`` (or an int seed) instead of a JAX threefry key, so the DRAWS
differ from the reference's while their distributions and ranges are the reference's (parity is defined on identical
inputs, not identical generators); and the procedural B-spline scar generator (deepx/utils_scars.py, 429 lines of
skimage/scipy) is replaced by a seeded blob field with the same output contract (smooth map in [0, 1], 1 = healthy).

``ensemble`` is what BASELINE config 4 runs: many independent random sequences stepped as ONE batched launch per
checkpoint segment (``generate_FKset.py:109-126`` loops over the seeds one after the other), ranks taking contiguous
ranges of seeds with no communication; snapshots leave through the asynchronous writer of ``cardiax_b200.io``.
"""
from functools import partial

import numpy as np
import torch

from . import convert, io, options, solve, stimulus


def _rng(rng):
    return rng if isinstance(rng, np.random.Generator) else np.random.default_rng(rng)


def _split(rng, n=2):
    """Independent child generators (the role of ``jax.random.split``)."""
    return _rng(rng).spawn(n)


def random_protocol(rng, min_start=0, max_start=1000, min_period=400, max_period=1e9):
    """deepx/generate.py:16-27 -- start ~ U{min_start..max_start-1}, duration 2, period ~ U{min_period..max_period-1}
    (shape-(1,) integer arrays, like the reference's)."""
    rng_1, rng_2 = _split(rng)
    start = rng_1.integers(int(min_start), int(max_start), (1,))
    duration = 2  # always instantaneous
    period = rng_2.integers(int(min_period), int(max_period), (1,))
    return stimulus.Protocol(start, duration, period)


def random_rectangular_stimulus(rng, shape, protocol, modulus=0.6):
    """deepx/generate.py:30-36."""
    rng_1, rng_2 = _split(rng)
    size = rng_2.integers(max(shape[0] // 100, 10), shape[0] // 3, (2,))
    centre = rng_1.integers(int(size.min()), shape[0], (2,))
    return stimulus.rectangular(shape, centre, size, modulus, protocol)


def random_linear_stimulus(rng, shape, protocol, modulus=0.6):
    """deepx/generate.py:39-45 -- |N(0, 1)| * 0.2 coverage, direction ~ U{0, 1, 2}."""
    rng_1, _ = _split(rng)
    coverage = abs(float(rng_1.normal()))
    direction = int(rng_1.integers(0, 3))
    return stimulus.linear(shape, direction, coverage * 0.2, modulus, protocol)


def random_triangular_stimulus(rng, shape, protocol, modulus=0.6):
    """deepx/generate.py:48-56."""
    rng_1, _ = _split(rng)
    angle, coverage = np.abs(rng_1.normal(size=2))
    direction = int(rng_1.integers(0, 3))
    return stimulus.triangular(shape, direction, float(angle) * 45, float(coverage) * 0.2, modulus, protocol)


def random_stimulus(rng, shape, min_start=0, max_start=0):
    """deepx/generate.py:59-76 -- one of {rectangular, triangular, linear}, amplitude 20."""
    stimuli_fn = (random_rectangular_stimulus, random_triangular_stimulus, random_linear_stimulus)
    rng_1, rng_2, rng_3 = _split(rng, 3)
    protocol = random_protocol(rng_1, min_start=min_start, max_start=max_start)
    modulus = 20.0
    stimulus_fn = partial(stimuli_fn[int(rng_2.integers(0, len(stimuli_fn)))], shape=shape, protocol=protocol,
                          modulus=modulus)
    return stimulus_fn(rng_3)


def random_diffusivity_scar(rng, shape):
    """Stand-in for deepx/utils_scars.py:409-426: a few random elliptic scars, Gaussian-tapered, returned as
    ``1 - scar`` in [0, 1]."""
    from scipy import ndimage
    rng = _rng(rng)
    H, W = shape
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    scar = np.zeros(shape, np.float32)
    for _ in range(int(rng.integers(2, 6))):
        cy, cx = rng.uniform(0, H), rng.uniform(0, W)
        ry, rx = rng.uniform(0.03, 0.15) * H, rng.uniform(0.03, 0.15) * W
        th = rng.uniform(0, np.pi)
        a = (yy - cy) * np.cos(th) + (xx - cx) * np.sin(th)
        b = -(yy - cy) * np.sin(th) + (xx - cx) * np.cos(th)
        scar = np.maximum(scar, ((a / ry) ** 2 + (b / rx) ** 2 <= 1.0).astype(np.float32))
    sigma = max(1.0, abs(float(rng.normal())) * H / 40)
    scar = ndimage.gaussian_filter(scar, sigma, mode="nearest")
    return (1.0 - np.clip(scar, 0.0, 1.0)).astype(np.float32)


def random_diffusivity(rng, shape, domain=(0.0001, 0.001)):
    """deepx/generate.py:79-83."""
    c = random_diffusivity_scar(rng, shape)
    if float(c.max()) == float(c.min()):
        return np.full(shape, domain[1], np.float32)
    return convert.diffusivity_rescale(c, domain).astype(np.float32)


def _random_inputs(rng, params, shape, n_stimuli, stop, dt):
    """The drawing part of random_sequence (deepx/generate.py:101-116)."""
    rngs = _split(rng, n_stimuli)
    max_start = np.arange(1, convert.ms_to_units(stop, dt), convert.ms_to_units(solve._scalar(params.tau_d) * 1000, dt))
    stimuli = [random_stimulus(rngs[i], shape, min_start=max_start[i], max_start=max_start[i] + 1) for i in range(n_stimuli)]
    diffusivity = random_diffusivity(rngs[-1], shape)
    return stimuli, diffusivity


def random_sequence(rng, params, filepath, shape=(1200, 1200), n_stimuli=2, start=0, stop=1000, step=1, dt=0.01, dx=0.01,
                    reshape=None, use_memory=False, plot_while=True):
    """deepx/generate.py:86-132 -- times in ms; stimuli start ``tau_d * 1000`` ms apart."""
    stimuli, diffusivity = _random_inputs(rng, params, shape, n_stimuli, stop, dt)
    return sequence(start=convert.ms_to_units(start, dt), stop=convert.ms_to_units(stop, dt),
                    step=convert.ms_to_units(step, dt), dt=dt, dx=dx, params=params, diffusivity=diffusivity,
                    stimuli=stimuli, filename=filepath, reshape=reshape, use_memory=use_memory, plot_while=plot_while)


def sequence(start, stop, step, dt, dx, params, diffusivity, stimuli, filename, reshape=None, use_memory=False,
             plot_while=True):
    """deepx/generate.py:135-210 -- checkpointed run writing ``states (T, 3, H', W')``.  Snapshot resize, D2H copy and
    file write overlap the solver (``io.AsyncSnapshotWriter``); ``use_memory`` is accepted and makes no difference
    (the pinned ring plays that role)."""
    shape = tuple(diffusivity.shape)
    if options.verbose:
        print("Tissue size is: {} - Computing on grid {}".format(convert.shape_to_realsize(shape, dx), shape))
        print("Checkpointing every {} steps".format(step))
        print("Cell parameters", params)
    if plot_while:
        try:
            from . import plot
            plot.plot_diffusivity(diffusivity)
            plot.plot_stimuli(stimuli)
        except Exception:  # matplotlib is optional in this image
            pass
    return io.sequence(start, stop, step, dt, dx, params, diffusivity, stimuli, filename, reshape=reshape,
                       use_memory=use_memory, plot_while=False)


def shard(seeds, rank, world_size):
    """Contiguous share of ``seeds`` for ``rank``: ceil(n / world_size) each, the last ranks may get fewer (or none)."""
    seeds = list(seeds)
    per = (len(seeds) + world_size - 1) // world_size
    return seeds[rank * per:(rank + 1) * per]


class _Scatter:
    """``dset[t] = (3, batch, H', W')`` -> ``states`` dataset of every member file."""

    def __init__(self, dsets):
        self.dsets = dsets

    def __setitem__(self, t, arr):
        for b, d in enumerate(self.dsets):
            d[t] = arr[:, b]


def ensemble(seeds, params, filepattern, shape=(256, 256), n_stimuli=3, start=0, stop=1000, step=1, dt=0.01, dx=0.01,
             reshape=None, rank=0, world_size=1, chunk=128):
    """BASELINE config 4: ``random_sequence`` for every seed in ``seeds`` -- this rank's contiguous share of them --
    stepped ``chunk`` tissues at a time as one batch (no communication between ranks).  ``filepattern % seed`` names
    each member's file.  Returns the seeds this rank generated."""
    mine = shard(seeds, rank, world_size)
    dev = solve._device()
    out_shape = tuple(reshape) if reshape is not None else tuple(shape)
    cps = np.arange(convert.ms_to_units(start, dt), convert.ms_to_units(stop, dt), convert.ms_to_units(step, dt))
    for c0 in range(0, len(mine), chunk):
        members = mine[c0:c0 + chunk]
        drawn = [_random_inputs(s, params, shape, n_stimuli, stop, dt) for s in members]
        D = torch.as_tensor(np.stack([d for _, d in drawn])).to(dev)
        stim = [[stimulus.Stimulus(s.protocol, solve._as_f32(s.field, dev)) for s in ss] for ss, _ in drawn]
        files = []
        for seed, (ss, d) in zip(members, drawn):
            f = io.init(filepattern % seed, out_shape, n_iter=len(cps), n_stimuli=len(ss))
            io.add_params(f, params, d, dt, dx, shape=out_shape)
            io.add_stimuli(f, ss, shape=out_shape)
            io.add_diffusivity(f, d, shape=out_shape)
            files.append(f)
        nb = len(members)
        state = solve.State(*[x.unsqueeze(0).repeat(nb, 1, 1) for x in solve.init(shape)])
        writer = io.AsyncSnapshotWriter(_Scatter([f["states"] for f in files]), (3, nb) + out_shape)
        try:
            for i in range(len(cps) - 1):
                state = solve._forward_euler(state, cps[i], cps[i + 1], params, D, stim, dt, dx)
                writer.submit(state, i)
        finally:
            writer.close()
        for f in files:
            f.close()
    return mine
