"""Snapshot path of the reference (cardiax/io.py:13-124) with asynchronous device-to-host copies.

Layout kept (io.py:18-22, 31-35, 43, 53-63): ``states (T, 3, H', W') f32`` in (v, w, u) order, ``stimuli``,
``params/{D, dt, dx, <14 names>}``, ``diffusivity``, ``field/start/duration/period``.

* ``init``/``add_params``/``add_diffusivity``/``add_stimuli``/``add_state``/``add_states``/``load*``/``imresize`` keep the
  reference's names and arguments.  ``h5py`` is optional: without it the same calls go to an in-memory store that is
  written as one ``.npz`` on ``close()`` (dataset names with ``/`` kept as keys).
* ``imresize`` is ``jax.image.resize(a, shape, "bilinear")`` (anti-aliased triangle filter, half-pixel centres, edge
  weights renormalised) evaluated on the GPU by ``fk_resize_kernel`` (one thread per output pixel, the separable filter
  taps staged in shared memory; all planes of a snapshot in one launch).
* ``AsyncSnapshotWriter`` is what the north star asks for: snapshots are (optionally resized and) copied into a ring of
  PINNED host buffers on a SIDE stream, ordered after the producing kernels by an event, and written to the dataset by
  a host thread -- the solver stream never waits for the disk or the PCIe copy.
"""
import os
import queue
import threading

import numpy as np
import torch

from .params import Params
from .stimulus import Protocol, Stimulus

try:  # optional
    import h5py  # noqa: F401
    _HAVE_H5 = True
except Exception:  # pragma: no cover - depends on the image
    h5py = None
    _HAVE_H5 = False


# --------------------------------------------------------------------------- storage without h5py
class _Dataset:
    def __init__(self, array):
        self.array = array

    def __setitem__(self, key, value):
        self.array[key] = _to_numpy(value)

    def __getitem__(self, key):
        return self.array[key]

    def __len__(self):
        return len(self.array)

    @property
    def shape(self):
        return self.array.shape


class NpzStore:
    """Minimal stand-in for the part of ``h5py.File`` the reference uses; saved as ``<path>.npz`` on close.  Datasets
    created by shape above ``MEMMAP_BYTES`` (the ``states`` of a long run: 1000 snapshots of 3 x 256^2 are 786 MB, and
    an ensemble holds one per member) live in disk-backed ``<path>.<name>.npy`` memory maps next to the ``.npz``, which
    then only lists them -- like HDF5, the run's snapshots never have to fit in host memory."""
    MEMMAP_BYTES = 64 << 20

    def __init__(self, path, mode="w"):
        self.path = path if path.endswith(".npz") else path + ".npz"
        self.data = {}
        self.external = {}
        if mode == "r":
            with np.load(self.path, allow_pickle=False) as f:
                for k in f.files:
                    if k == "__external__":
                        for item in f[k]:
                            name, fname = str(item).split("=", 1)
                            self.data[name] = _Dataset(np.load(os.path.join(os.path.dirname(self.path), fname), mmap_mode="r"))
                    else:
                        self.data[k] = _Dataset(f[k])

    def create_dataset(self, name, shape=None, dtype=None, data=None):
        if data is not None:
            arr = np.array(_to_numpy(data), dtype=dtype)
        else:
            dt = np.dtype(dtype or "float32")
            if int(np.prod(shape, dtype=np.int64)) * dt.itemsize > self.MEMMAP_BYTES:
                fname = os.path.basename(self.path)[:-4] + "." + name.replace("/", "_") + ".npy"
                arr = np.lib.format.open_memmap(os.path.join(os.path.dirname(self.path), fname), mode="w+", dtype=dt,
                                                shape=tuple(int(x) for x in shape))
                self.external[name] = fname
            else:
                arr = np.zeros(shape, dtype=dt)
        self.data[name] = _Dataset(arr)
        return self.data[name]

    def __contains__(self, name):
        return name in self.data

    def __getitem__(self, name):
        if name in self.data:
            return self.data[name]
        sub = {k[len(name) + 1:]: v for k, v in self.data.items() if k.startswith(name + "/")}
        if not sub:
            raise KeyError(name)
        return sub

    def __iter__(self):
        return iter(self.data)

    def close(self):
        small = {k: v.array for k, v in self.data.items() if k not in self.external}
        for k in self.external:
            self.data[k].array.flush()
        if self.external:
            small["__external__"] = np.array(["%s=%s" % kv for kv in sorted(self.external.items())])
        np.savez(self.path, **small)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass


def _to_numpy(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    if isinstance(x, (list, tuple)) and len(x) and isinstance(x[0], torch.Tensor):
        return np.stack([t.detach().cpu().numpy() for t in x])
    return np.asarray(x)


def _open(path, mode):
    if _HAVE_H5:
        return h5py.File(path, mode)
    return NpzStore(path, mode)


# --------------------------------------------------------------------------- reference API
def init(path, shape, n_iter, n_stimuli, n_variables=3):
    """io.py:13-23."""
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    hdf5 = _open(path, "w")
    if "states" not in hdf5:
        hdf5.create_dataset("states", shape=(n_iter, n_variables, *shape), dtype="float32")
    if "stimuli" not in hdf5:
        hdf5.create_dataset("stimuli", shape=(n_stimuli, *shape), dtype="float32")
    return hdf5


def add_params(hdf5, params, diffusivity, dt, dx, shape=None):
    """io.py:26-36."""
    if shape is not None:
        diffusivity = imresize(diffusivity, shape)
    hdf5.create_dataset("params/D", data=_to_numpy(diffusivity))
    hdf5.create_dataset("params/dt", data=dt)
    hdf5.create_dataset("params/dx", data=dx)
    for i in range(len(params)):
        hdf5.create_dataset("params/" + params._fields[i], data=params[i])
    return True


def add_diffusivity(hdf5, diffusivity, shape=None):
    """io.py:39-44."""
    if shape is not None:
        diffusivity = imresize(diffusivity, shape)
    hdf5.create_dataset("diffusivity", data=_to_numpy(diffusivity))
    return True


def add_stimuli(hdf5, stimuli, shape=None):
    """io.py:47-64."""
    if shape is not None:
        fields = [_to_numpy(imresize(s.field, shape)) for s in stimuli]
    else:
        fields = [_to_numpy(s.field) for s in stimuli]
    hdf5.create_dataset("field", data=np.array(fields, dtype=np.float32))
    scal = lambda x: np.asarray(_to_numpy(x)).reshape(-1)[0]  # noqa: E731
    hdf5.create_dataset("start", data=[scal(s.protocol.start) for s in stimuli])
    hdf5.create_dataset("duration", data=[scal(s.protocol.duration) for s in stimuli])
    hdf5.create_dataset("period", data=[scal(s.protocol.period) for s in stimuli])
    return True


def add_state(dset, state, t, shape=None):
    """io.py:67-72 -- synchronous (resize +) store of one snapshot; see AsyncSnapshotWriter for the overlapped path."""
    if shape is not None:
        parts = tuple(state)
        if all(isinstance(x, torch.Tensor) and x.is_cuda and x.dim() == 2 for x in parts) and \
                tuple(parts[0].shape) != tuple(shape[-2:]):
            H, W = parts[0].shape      # v, w, u resized into one packed (3, H', W') array by ONE launch, no stack
            state = _resize_planes([x.to(torch.float32).contiguous() for x in parts], H, W, tuple(shape[-2:]))
        else:
            state = imresize(torch.stack([torch.as_tensor(np.asarray(x)) if not isinstance(x, torch.Tensor) else x
                                          for x in parts]), tuple(shape[-2:]))
    dset[t] = _to_numpy(state)
    return True


def add_states(dset, states, start, end):
    """io.py:75-78."""
    dset[start:end] = np.stack([_to_numpy(s) for s in states]) if isinstance(states, (list, tuple)) else _to_numpy(states)
    return True


def load(path, start=None, end=None, step=None):
    """io.py:81-83."""
    f = _open(path, "r")
    try:
        return [f[dset][start:end:step] for dset in f]
    finally:
        if _HAVE_H5:
            f.close()


def load_state(dset, start, end, step):
    return dset[start:end:step]


def load_stimuli(file):
    """io.py:90-96."""
    stimuli = []
    for i in range(len(file["field"])):
        protocol = Protocol(file["start"][i], file["duration"][i], file["period"][i])
        stimuli.append(Stimulus(protocol, file["field"][i]))
    return stimuli


def load_params(filepath):
    """io.py:99-110."""
    params = {}
    D = None
    f = _open(filepath, "r")
    stored = f["params"]
    for key in stored:
        if key == "D":
            D = stored[key][...]
        else:
            params[key] = stored[key][...]
    if _HAVE_H5:
        f.close()
    return Params(*[params[k] for k in Params._fields]), D


def load_diffusivity(filepath):
    f = _open(filepath, "r")
    try:
        return f["diffusivity"][:]
    finally:
        if _HAVE_H5:
            f.close()


# --------------------------------------------------------------------------- jax.image.resize(..., "bilinear")
_resize_ws = {}   # (device, stream, H, W, Ho, Wo, n) -> workspace holding the uploaded filter tables


def _resize_planes(planes, H, W, size, out=None):
    """``planes``: list of contiguous fp32 CUDA (H, W) tensors -> packed (len(planes), H', W') tensor, ONE launch of
    ``fk_resize_kernel`` on the current stream (filter tables built by the library -- csrc/fk_aux.h -- and kept in a
    per-geometry workspace, so a repeated snapshot shape costs the kernel launch only)."""
    import ctypes
    from . import _lib
    L = _lib.lib()
    dev = planes[0].device
    Ho, Wo = int(size[0]), int(size[1])
    n = len(planes)
    if out is None:
        out = torch.empty((n, Ho, Wo), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev)
    key = (str(dev), stream.cuda_stream, H, W, Ho, Wo, n)   # per stream: the tables are uploaded in stream order
    ws = _resize_ws.get(key)
    ready = ws is not None
    if not ready:
        ws = torch.empty(L.fk_resize_workspace_bytes(H, W, Ho, Wo, n), dtype=torch.uint8, device=dev)
    ptrs = (ctypes.c_void_p * n)(*[p.data_ptr() for p in planes])
    _lib.check(L.fk_resize_bilinear(ptrs, n, H, W, out.data_ptr(), Ho, Wo, ws.data_ptr(), ws.numel(), int(ready),
                                    ctypes.c_void_p(stream.cuda_stream)))
    if not ready:   # remembered only once the call that uploads the filter tables has succeeded
        if len(_resize_ws) > 32:
            _resize_ws.clear()
        _resize_ws[key] = ws
    return out


def imresize(a, size, method="bilinear"):
    """io.py:118-124 -- ``jax.image.resize(a, a.shape[:-2] + size, "bilinear")`` of a 2-D or 3-D (or deeper) array:
    anti-aliased triangle filter, half-pixel centres.  Runs on the GPU (host arrays are uploaded; no CPU fallback)."""
    if method != "bilinear":
        raise NotImplementedError("only the reference's default 'bilinear' is provided")
    from . import solve
    if not isinstance(a, torch.Tensor):
        a = torch.as_tensor(np.asarray(a))
    H, W = a.shape[-2:]
    if (H, W) == tuple(size):
        return a.to(torch.float32).clone()
    a = solve._as_f32(a)
    lead = tuple(a.shape[:-2])
    flat = a.reshape((-1, H, W))
    out = _resize_planes([flat[i] for i in range(flat.shape[0])], H, W, size)
    return out.reshape(lead + (int(size[0]), int(size[1])))


# --------------------------------------------------------------------------- asynchronous snapshots
class AsyncSnapshotWriter:
    """Overlapped snapshotting: ``submit(state, t)`` returns at once; the copy runs on a side stream into pinned memory
    and a host thread stores it into ``dset[t]``.

    dset        any object with ``__setitem__`` (the ``states`` dataset of ``init``)
    shape       snapshot shape (3, H', W'); if (H', W') differs from the state's grid it is resized on the GPU first
    slots       pinned ring depth; ``submit`` blocks only when all slots are still waiting for the writer thread
    """

    def __init__(self, dset, shape, slots=4, device=None):
        self.dset = dset
        self.shape = tuple(shape)
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.stream = torch.cuda.Stream(device=self.device)
        self.pinned = [torch.empty(self.shape, dtype=torch.float32).pin_memory() for _ in range(slots)]
        self.free = queue.Queue()
        for i in range(slots):
            self.free.put(i)
        self.work = queue.Queue()
        self.error = None
        self.bytes_copied = 0
        self.thread = threading.Thread(target=self._drain, daemon=True)
        self.thread.start()

    def submit(self, state, t):
        slot = self.free.get()
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))      # after the kernels that produced `state`
        keep = tuple(state)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            H, W = keep[0].shape[-2:]
            if (H, W) != self.shape[-2:]:   # v, w, u (of every tissue of a batch) in one launch
                planes = [p for x in keep for p in x.contiguous().reshape(-1, H, W)]
                arr = _resize_planes(planes, H, W, self.shape[-2:]).reshape(self.shape)
            else:
                arr = torch.stack(keep)
            self.pinned[slot].copy_(arr, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.stream)
        for x in keep:                                              # the side stream still reads them
            x.record_stream(self.stream)
        self.bytes_copied += self.pinned[slot].numel() * 4
        self.work.put((slot, t, done))

    def _drain(self):
        while True:
            item = self.work.get()
            if item is None:
                return
            slot, t, done = item
            try:
                done.synchronize()
                self.dset[t] = self.pinned[slot].numpy()
            except Exception as e:  # noqa: BLE001 - reported by close()
                self.error = e
            self.free.put(slot)

    def close(self):
        self.work.put(None)
        self.thread.join()
        if self.error is not None:
            raise self.error


def sequence(start, stop, step, dt, dx, params, diffusivity, stimuli, filename, reshape=None, use_memory=False,
             plot_while=False):
    """deepx/generate.py:135-210 -- checkpointed run with snapshots; D2H and file writes overlap the solver."""
    from . import solve
    shape = tuple(diffusivity.shape)
    out_shape = tuple(reshape) if reshape is not None else shape
    checkpoints = np.arange(int(start), int(stop), int(step))
    hdf5 = init(filename, out_shape, n_iter=len(checkpoints), n_stimuli=len(stimuli))
    add_params(hdf5, params, diffusivity, dt, dx, shape=out_shape)
    add_stimuli(hdf5, stimuli, shape=out_shape)
    add_diffusivity(hdf5, diffusivity, shape=out_shape)
    states_dset = hdf5["states"]
    state = solve.init(shape)
    # the static inputs go to the device ONCE: every segment then sees the same tensor objects (no re-upload, and the
    # solver's per-tensor uniform-diffusivity verdict is reused)
    D_dev = solve._as_f32(diffusivity)
    stim_dev = [type(s)(s.protocol, solve._as_f32(s.field)) for s in stimuli]
    writer = None
    try:
        writer = AsyncSnapshotWriter(states_dset, (3,) + out_shape)
        for i in range(len(checkpoints) - 1):
            state = solve._forward_euler(state, checkpoints[i], checkpoints[i + 1], params, D_dev, stim_dev, dt, dx)
            writer.submit(state, i)
    finally:
        try:
            if writer is not None:
                writer.close()
        finally:
            hdf5.close()
    return state
