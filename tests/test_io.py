"""Snapshot path (cardiax_b200/io.py): storage layout, bilinear anti-aliased resize, async writer."""
import os

import numpy as np
import pytest
import torch

import oracle as O


def _jax_resize_bilinear_numpy(a, size):
    """Direct (slow, fp64) statement of jax.image.resize(a, shape, 'bilinear') with antialias=True."""
    def mat(n_in, n_out):
        inv = n_in / n_out
        ks = max(inv, 1.0)
        out = np.zeros((n_in, n_out))
        for j in range(n_out):
            sf = (j + 0.5) * inv - 0.5
            w = np.array([max(0.0, 1.0 - abs(sf - i) / ks) for i in range(n_in)])
            tot = w.sum()
            w = w / tot if abs(tot) > 1000 * np.finfo(np.float32).eps else 0 * w
            if not (-0.5 <= sf <= n_in - 0.5):
                w = 0 * w
            out[:, j] = w
        return out
    return np.einsum("...hw,hi,wj->...ij", np.asarray(a, np.float64), mat(a.shape[-2], size[0]), mat(a.shape[-1], size[1]))


def test_imresize_matches_the_jax_algorithm():
    from cardiax_b200 import io
    rng = np.random.default_rng(0)
    for shape, size in (((3, 24, 36), (8, 9)), ((20, 20), (7, 13)), ((12, 12), (24, 30)), ((3, 10, 10), (10, 10))):
        a = rng.random(shape).astype(np.float32)
        got = io.imresize(torch.as_tensor(a), size).numpy()
        assert got.shape == shape[:-2] + size
        assert np.abs(got - _jax_resize_bilinear_numpy(a, size)).max() < 1e-5
    # constant stays constant; exact 2x box average of a ramp
    assert np.allclose(io.imresize(torch.full((16, 16), 3.0), (5, 7)).numpy(), 3.0, atol=1e-6)
    ramp = torch.arange(16, dtype=torch.float32).repeat(16, 1)
    assert np.allclose(io.imresize(ramp, (16, 4)).numpy()[0][1:3], [5.5, 9.5], atol=1e-5)


def test_storage_layout_round_trip(tmp_path):
    from cardiax_b200 import io, params, stimulus
    shape, out = (20, 30), (10, 15)
    path = os.path.join(tmp_path, "run", "seq.hdf5")
    D = torch.full(shape, 1e-3)
    stim = [stimulus.Stimulus(stimulus.Protocol(0, 2, 1e9), torch.ones(shape)),
            stimulus.Stimulus(stimulus.Protocol(np.array([40]), 2, np.array([400])), torch.zeros(shape))]
    f = io.init(path, out, n_iter=4, n_stimuli=2)
    io.add_params(f, params.PARAMSET_3, D, 0.01, 0.02, shape=out)
    io.add_stimuli(f, stim, shape=out)
    io.add_diffusivity(f, D, shape=out)
    st = O.init(shape)
    io.add_state(f["states"], [torch.as_tensor(x) for x in st], 1, shape=(3, *out))
    io.add_states(f["states"], [np.full((3, *out), 7.0, np.float32)] * 2, 2, 4)
    f.close()
    p, Dl = io.load_params(path)
    assert tuple(float(x) for x in p) == tuple(float(x) for x in params.PARAMSET_3) and Dl.shape == out
    assert io.load_diffusivity(path).shape == out
    g = io._open(path, "r")
    assert g["states"].shape == (4, 3, *out)
    assert np.allclose(g["states"][1][0], 1.0) and np.allclose(g["states"][1][2], 0.0) and np.allclose(g["states"][3], 7.0)
    ls = io.load_stimuli(g)
    assert len(ls) == 2 and float(ls[1].protocol.start) == 40 and np.allclose(ls[0].field, 1.0)


@pytest.mark.gpu
def test_async_writer_equals_synchronous_snapshots(tmp_path):
    from cardiax_b200 import io, options, solve
    options.verbose = False
    shape, out = (96, 128), (24, 32)
    st = solve.State(*[torch.rand(shape, device="cuda") for _ in range(3)])
    D = torch.full(shape, 1e-3, device="cuda")
    sync = np.zeros((6, 3, *out), np.float32)
    asyn = np.zeros((6, 3, *out), np.float32)
    w = io.AsyncSnapshotWriter(asyn, (3, *out), slots=2)
    s = st
    for i in range(6):
        s = solve._forward_euler(s, i * 5, i * 5 + 5, O.PARAMSETS["3"], D, [], 0.01, 0.01)
        w.submit(s, i)
        io.add_state(sync, s, i, shape=(3, *out))
    w.close()
    assert np.array_equal(sync, asyn) and w.bytes_copied == 6 * 3 * out[0] * out[1] * 4
    # the reference's data-generation driver on top of it
    final = io.sequence(0, 40, 10, 0.01, 0.01, O.PARAMSETS["3"], D, [], os.path.join(tmp_path, "s.hdf5"), reshape=out)
    g = io._open(os.path.join(tmp_path, "s.hdf5"), "r")
    ref = solve._forward_euler(solve.init(shape), 0, 30, O.PARAMSETS["3"], D, [], 0.01, 0.01)
    assert all(torch.equal(a, b) for a, b in zip(final, ref))
    assert np.allclose(g["states"][2], io.imresize(torch.stack(tuple(ref)), out).cpu().numpy())
