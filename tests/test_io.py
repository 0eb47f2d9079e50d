"""Snapshot path (cardiax_b200/io.py): storage layout, bilinear anti-aliased resize, async writer."""
import os

import numpy as np
import pytest
import torch

import oracle as O
from oracle import fk_oracle_ext as X


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


RESIZE_CASES = (((3, 24, 36), (8, 9)), ((20, 20), (7, 13)), ((12, 12), (24, 30)), ((3, 10, 10), (10, 10)),
                ((2, 3, 75, 61), (16, 13)), ((5, 5), (1, 1)), ((1, 7), (3, 2)), ((3, 150, 150), (32, 32)))


def _check_resize(fn):
    rng = np.random.default_rng(0)
    for shape, size in RESIZE_CASES:
        a = rng.random(shape).astype(np.float32)
        got = fn(a, size)
        assert got.shape == shape[:-2] + size
        # fp32 accumulation of O(1) values: tolerance 1e-5 absolute (observed < 5e-7)
        assert np.abs(got - _jax_resize_bilinear_numpy(a, size)).max() < 1e-5, (shape, size)
        assert np.abs(got - X.resize_bilinear(a, size)).max() < 1e-5, (shape, size)
    # constant stays constant; exact 2x box average of a ramp
    assert np.allclose(fn(np.full((16, 16), 3.0, np.float32), (5, 7)), 3.0, atol=1e-6)
    ramp = np.tile(np.arange(16, dtype=np.float32), (16, 1))
    assert np.allclose(fn(ramp, (16, 4))[0][1:3], [5.5, 9.5], atol=1e-5)


def test_resize_oracle_is_the_jax_algorithm():
    rng = np.random.default_rng(1)
    for shape, size in RESIZE_CASES:
        a = rng.random(shape).astype(np.float32)
        assert np.abs(X.resize_bilinear(a, size) - _jax_resize_bilinear_numpy(a, size)).max() < 1e-12


def test_resize_kernel_body_on_the_cpu_matches_the_oracle():
    from tests.emu import emu
    _check_resize(emu.resize)


@pytest.mark.gpu
def test_imresize_matches_the_jax_algorithm():
    from cardiax_b200 import _lib, io
    before = _lib.lib().fk_launch_count()
    _check_resize(lambda a, size: io.imresize(torch.as_tensor(a), size).cpu().numpy())
    assert _lib.lib().fk_launch_count() > before          # the library's kernel ran, not a torch op
    big = torch.rand((3, 1200, 1200), device="cuda")       # the reference's standard snapshot: 1200^2 -> 256^2
    got = io.imresize(big, (256, 256)).cpu().numpy()
    assert np.abs(got - X.resize_bilinear(big.cpu().numpy(), (256, 256))).max() < 1e-5


@pytest.mark.gpu
def test_storage_layout_round_trip_resized(tmp_path):
    _storage_round_trip(tmp_path, (20, 30), (10, 15))


def test_storage_layout_round_trip(tmp_path):
    _storage_round_trip(tmp_path, (20, 30), (20, 30))      # resizing runs on the GPU only: same shape on the CPU


def _storage_round_trip(tmp_path, shape, out):
    from cardiax_b200 import io, params, stimulus
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


def test_npz_store_keeps_large_datasets_on_disk(tmp_path, monkeypatch):
    """Without h5py the snapshot store must not hold a long run (or an ensemble's worth of them) in host memory: datasets
    above the threshold are disk-backed memory maps next to the .npz and read back lazily."""
    from cardiax_b200 import io
    monkeypatch.setattr(io, "_HAVE_H5", False)
    monkeypatch.setattr(io.NpzStore, "MEMMAP_BYTES", 1024)
    path = os.path.join(tmp_path, "run", "big.hdf5")
    f = io.init(path, (16, 16), n_iter=5, n_stimuli=1)          # states: 5 * 3 * 256 * 4 B > 1 KiB -> memory map
    assert isinstance(f["states"].array, np.memmap) and not isinstance(f["stimuli"].array, np.memmap)
    io.add_state(f["states"], [np.full((16, 16), float(i + 1), np.float32) for i in range(3)], 2)
    io.add_diffusivity(f, np.full((16, 16), 1e-3, np.float32))
    f.close()
    assert os.path.exists(os.path.join(tmp_path, "run", "big.hdf5.states.npy"))
    g = io._open(path, "r")
    assert g["states"].shape == (5, 3, 16, 16) and np.all(g["states"][2][1] == 2.0) and np.all(g["states"][0] == 0.0)
    assert np.allclose(io.load_diffusivity(path), 1e-3)
