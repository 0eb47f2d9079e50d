"""Row-slab decomposition of ONE large tissue over the GPUs of a node (BASELINE config 5).

Not in the reference (it runs a tissue on a single device); SURVEY 8e defines it.  One process per GPU.  Rank r owns a
contiguous block of rows and keeps ``halo_launches * 4T`` (default 8 * 8 = 64) halo rows of u, v, w from each neighbour,
so halos are exchanged once per GROUP of ``halo_launches`` kernel launches (T Euler steps each); between exchanges a
launch also recomputes the shrinking apron of halo rows it still needs.

The exchange is fused into the step kernel (``comm="peer"``, the default on CUDA): the state lives in buffers allocated
with ``fk_peer_alloc`` (cudaMalloc + CUDA IPC) that the two neighbouring ranks map into their address space, and the
LAST launch of a group -- the ``FK_STORE_MIRROR`` instantiation of the streaming kernel, ``fk_euler_rows_peer`` -- stores
the rows a neighbour needs straight into that neighbour's halo over NVLink while it produces them, so the transfer
overlaps the arithmetic row by row and no NCCL kernel competes with the single-wave step kernel for the SMs.  A
sequence number written into a flag in the neighbour's memory after the launch (``fk_peer_signal``) releases the
neighbour's next group, which waits for it with a stream memory operation (``fk_peer_wait``): no host thread, no SM.

Three state buffers make this safe without any reverse handshake: X[0], X[1] (peer visible) and Y (private).  Group g
starts from X[g % 2], ping-pongs between it and Y, and its last launch writes own rows into X[(g + 1) % 2] while the
neighbours' last launches write that buffer's halo rows.  A neighbour can only be one group ahead (it needs this rank's
rows of group g to start group g + 1), and what it then writes, X[g % 2]'s halo, this rank stopped reading before it
signalled group g.

``comm="dist"`` is the plain ``torch.distributed`` send/recv exchange after the group's last launch (gloo on the CPU,
where the tests inject the emulated kernels as the backend; NCCL if asked for on CUDA).

The state stays in these buffers between ``advance`` calls: ``advance(None, t0, t1)`` continues from where the last call
ended (halos already valid), ``advance(state, t0, t1)`` loads new own rows first.  Physical tissue edges exist only at
the top of rank 0 and at the bottom of the last rank; left/right edges on every rank.  Each launch is the same
streaming kernel as the single-GPU path restricted to a window of output rows, so results are bit-identical to a
single-GPU run of the whole tissue (tests/test_slab_gloo.py: CPU, emulated kernels; tests/test_gpu_multi.py: GPUs).
"""
import ctypes

import numpy as np
import torch

from . import _lib, options
from .solve import State, _as_f32, _is_int, _params_struct, _scalar, _stimulus_struct


class CudaBackend:
    """fk_euler_rows(_peer) / fk_diffusivity_gradients on the current CUDA device and stream."""

    def __init__(self):
        self.L = _lib.lib()
        self.device = torch.device("cuda", torch.cuda.current_device())
        self._ws = torch.empty(4096, dtype=torch.uint8, device=self.device)
        self.fused_mirrors = 0          # launches whose halo mirror was done by the streaming kernel itself

    def dgrad(self, D, dx, phys_top, phys_bottom):
        DX, DY = torch.empty_like(D), torch.empty_like(D)
        H, W = D.shape
        _lib.check(self.L.fk_diffusivity_gradients(D.data_ptr(), DX.data_ptr(), DY.data_ptr(), H, W, 1,
                                                   np.float32(dx), int(phys_top), int(phys_bottom), self._stream()))
        return DX, DY

    @staticmethod
    def _stream():
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def euler_rows(self, src, dst, D, DX, DY, params, stimuli, t0, nsteps, dt, dx, phys_top, phys_bottom, row0, row1,
                   uniform, mirror=None):
        H, W = D.shape
        P = _params_struct(params)
        arr = (_lib.FkStimulus * max(1, len(stimuli)))()
        for i, s in enumerate(stimuli):
            arr[i] = _stimulus_struct(s.field.data_ptr(), s.protocol)
        o = _lib.FkOptions()
        self.L.fk_default_options(ctypes.byref(o))
        o.exact = int(options.numerics == "exact")
        o.kernel = int(options.kernel)
        o.cta_threads, o.rows_per_cta = int(options.cta_threads), int(options.rows_per_cta)
        o.phys_top, o.phys_bottom = int(phys_top), int(phys_bottom)
        o.uniform_diffusivity = int(uniform)
        o.safe_division = int(options.safe_division)
        o.counter_is_int = int(_is_int(t0))
        fused = ctypes.c_int(0)
        _lib.check(self.L.fk_euler_rows_peer(
            src[0].data_ptr(), src[1].data_ptr(), src[2].data_ptr(), dst[0].data_ptr(), dst[1].data_ptr(),
            dst[2].data_ptr(), D.data_ptr(), DX.data_ptr(), DY.data_ptr(), H, W, ctypes.byref(P), arr, len(stimuli),
            float(_scalar(t0)), int(nsteps), np.float32(dt), np.float32(dx), ctypes.byref(o), int(row0), int(row1),
            self._ws.data_ptr(), self._ws.numel(), self._stream(), ctypes.byref(mirror) if mirror is not None else None,
            ctypes.byref(fused)))
        self.fused_mirrors += fused.value


class _RawCuda:
    """A device allocation that is not torch's, presented through the CUDA array interface (zero-copy tensor view)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2,
                                         "strides": None}


class DistComm:
    """Halo exchange with torch.distributed point-to-point operations (gloo on CPU tensors, NCCL on CUDA tensors)."""
    kind = "dist"

    def __init__(self, runner):
        self.r = runner

    def alloc(self, Hb, W, dev):
        mk = lambda: [torch.zeros((Hb, W), dtype=torch.float32, device=dev) for _ in range(3)]  # noqa: E731
        return [mk(), mk()], mk()

    def mirror(self, parity):
        return None

    def wait(self):
        pass

    def after_last_launch(self, arrays, parity):
        self.exchange(arrays)

    def exchange(self, arrays):
        r, d = self.r, self.r.dist
        cpu_stage = arrays[0].is_cuda and d.get_backend(r.group) == "gloo"     # gloo moves host memory only
        ops, recvs = [], []
        for a in arrays:
            for nb, own, halo in ((r.rank - 1, slice(r.o0, r.o0 + r.Hh), slice(r.o0 - r.Hh, r.o0)),
                                  (r.rank + 1, slice(r.o1 - r.Hh, r.o1), slice(r.o1, r.o1 + r.Hh))):
                if nb < 0 or nb >= r.world:
                    continue
                if cpu_stage:
                    send, recv = a[own].cpu(), torch.empty((r.Hh, r.W), dtype=torch.float32)
                    recvs.append((a, halo, recv))
                else:
                    send, recv = a[own], a[halo]
                ops.append(d.P2POp(d.isend, send, nb, r.group))
                ops.append(d.P2POp(d.irecv, recv, nb, r.group))
        for w in (d.batch_isend_irecv(ops) if ops else []):
            w.wait()
        for a, halo, recv in recvs:
            a[halo] = recv.to(a.device)

    def close(self):
        pass


class PeerComm:
    """Halo exchange through CUDA-IPC-mapped peer memory: mirror stores by the step kernel + flags (module docstring)."""
    kind = "peer"

    def __init__(self, runner):
        self.r = runner
        self.L = _lib.lib()
        self.base = None
        self.mapped = {}
        self.seq = 0           # exchanges this rank has signalled == exchanges it expects from each neighbour

    def alloc(self, Hb, W, dev):
        r = self.r
        self.plane = Hb * W * 4
        nbytes = 6 * self.plane + 256
        base, handle = ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
        _lib.check(self.L.fk_peer_alloc(nbytes, ctypes.byref(base), handle))
        self.base = base.value
        X = [[torch.as_tensor(_RawCuda(self.base + (p * 3 + a) * self.plane, (Hb, W)), device=dev) for a in range(3)]
             for p in range(2)]
        for p in range(2):
            for a in range(3):
                if X[p][a].data_ptr() != self.base + (p * 3 + a) * self.plane:
                    raise RuntimeError("the CUDA array interface view of the exchange buffer is a copy, not an alias")
        self.flags = self.base + 6 * self.plane       # [0]: written by the rank above, [1]: by the rank below
        Y = [torch.zeros((Hb, W), dtype=torch.float32, device=dev) for _ in range(3)]
        # tell the neighbours where this rank's buffers are and how its rows are laid out
        mine = dict(handle=bytes(handle), o0=r.o0, o1=r.o1, Hb=Hb, dev=torch.cuda.current_device(), pid=__import__("os").getpid())
        info = [None] * r.world
        r.dist.all_gather_object(info, mine, group=r.group)
        self.nb = {}
        for side, nbr in ((0, r.rank - 1), (1, r.rank + 1)):
            if nbr < 0 or nbr >= r.world:
                continue
            ptr = ctypes.c_void_p()
            h = (ctypes.c_ubyte * 64).from_buffer_copy(info[nbr]["handle"])
            _lib.check(self.L.fk_peer_open(h, ctypes.byref(ptr)))
            self.mapped[side] = ptr.value
            i = info[nbr]
            # my top band -> the upper neighbour's bottom halo; my bottom band -> the lower neighbour's top halo
            dst_row0 = i["o1"] if side == 0 else i["o0"] - r.Hh
            self.nb[side] = dict(base=ptr.value, plane=i["Hb"] * W * 4, dst_row0=dst_row0,
                                 flag=ptr.value + 6 * i["Hb"] * W * 4 + 4 * (1 - side))
        r.dist.barrier(group=r.group)     # everybody has mapped everybody before anything is written
        return X, Y

    def mirror(self, parity):
        """FkPeerMirror for a launch whose result goes to X[parity]: own edge rows -> the neighbours' X[parity] halos."""
        r = self.r
        m = _lib.FkPeerMirror()
        for side, (row0, row1) in ((0, (r.o0, r.o0 + r.Hh)), (1, (r.o1 - r.Hh, r.o1))):
            nb = self.nb.get(side)
            if nb is None:
                continue
            for a, arr in enumerate((m.v, m.w, m.u)):          # buffers hold v, w, u in that order
                arr[side] = nb["base"] + (parity * 3 + a) * nb["plane"]
            m.row0[side], m.row1[side], m.dst_row0[side] = row0, row1, nb["dst_row0"]
        return m

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def signal(self):
        self.seq += 1
        for nb in self.nb.values():
            _lib.check(self.L.fk_peer_signal(ctypes.c_void_p(nb["flag"]), self.seq, self._stream()))

    def wait(self):
        """Order the current stream after the arrival of every exchange signalled so far (both neighbours)."""
        for side in self.nb:
            _lib.check(self.L.fk_peer_wait(ctypes.c_void_p(self.flags + 4 * side), self.seq, self._stream()))

    def after_last_launch(self, arrays, parity):
        self.signal()            # the launch mirrored its edge rows itself (or the library copied them): release the neighbours

    def exchange(self, arrays, parity):
        """Own edge rows of X[parity] -> the neighbours' halos by plain peer copies (after ``load``)."""
        r = self.r
        for side, (row0, row1) in ((0, (r.o0, r.o0 + r.Hh)), (1, (r.o1 - r.Hh, r.o1))):
            nb = self.nb.get(side)
            if nb is None:
                continue
            for a in range(3):
                dst = nb["base"] + (parity * 3 + a) * nb["plane"] + nb["dst_row0"] * r.W * 4
                _lib.check(self.L.fk_peer_copy(ctypes.c_void_p(dst), ctypes.c_void_p(arrays[a].data_ptr() + row0 * r.W * 4),
                                               (row1 - row0) * r.W * 4, self._stream()))
        self.signal()

    def close(self):
        if self.base is None:
            return
        torch.cuda.synchronize()
        try:
            self.r.dist.barrier(group=self.r.group)     # nobody unmaps memory a neighbour may still be writing
        except Exception:
            pass
        for ptr in self.mapped.values():
            self.L.fk_peer_close(ctypes.c_void_p(ptr))
        self.mapped = {}
        self.L.fk_peer_free(ctypes.c_void_p(self.base))
        self.base = None


class SlabRunner:
    """Advance this rank's rows of a row-decomposed tissue.

    state, diffusivity, stimuli fields: this rank's OWN rows, shape (H_local, W).  All ranks must make the same sequence
    of ``advance`` / ``load`` calls.  ``backend`` is the compute backend (CUDA by default; the CPU tests inject the
    emulated kernels); ``group`` the torch.distributed process group; ``comm``: "peer" (default on CUDA with more than
    one rank), "dist", or None for the default.
    """

    def __init__(self, state, diffusivity, params, stimuli, dt, dx, rank, world, steps_per_launch=0, halo_launches=8,
                 backend=None, group=None, overlap=True, comm=None):
        import torch.distributed as dist
        self.dist = dist
        self.rank, self.world, self.group = rank, world, group
        self.be = backend or CudaBackend()
        self.dev = state[0].device if isinstance(state[0], torch.Tensor) else self.be.device
        self.T = int(steps_per_launch or options.steps_per_launch or 2)
        self.M = int(halo_launches)
        self.F = 4 * self.T
        self.Hh = self.M * self.F                     # halo rows kept from each neighbour
        self.params, self.dt, self.dx = params, _scalar(dt), _scalar(dx)
        self.top, self.bot = rank == 0, rank == world - 1   # physical edges
        u = _as_f32(state[2], self.dev)
        self.Hl, self.W = u.shape
        if world > 1 and self.Hl < 2 * self.Hh + 8:
            raise ValueError("slab of %d rows is too thin for a halo of %d rows" % (self.Hl, self.Hh))
        self.o0 = 0 if self.top else self.Hh          # first own row inside the buffer
        self.o1 = self.o0 + self.Hl
        self.Hb = self.o1 + (0 if self.bot else self.Hh)
        if comm is None:
            comm = "peer" if (isinstance(self.be, CudaBackend) and world > 1) else "dist"
        self.comm = PeerComm(self) if comm == "peer" else DistComm(self)
        self._setup = DistComm(self)                  # the one-time exchange of the static maps
        # static maps with their halos (exchanged once)
        self.D = self._with_halo(_as_f32(diffusivity, self.dev))
        self.uniform = self._uniform_everywhere(self.D)
        self.DX, self.DY = self.be.dgrad(self.D, self.dx, self.top, self.bot)
        self.stimuli = [type(s)(s.protocol, self._with_halo(_as_f32(s.field, self.dev))) for s in stimuli]
        self.X, self.Y = self.comm.alloc(self.Hb, self.W, self.dev)
        self.par = 0                 # the state (own rows + valid halos) is in X[par]
        self.exchanges = 0
        self.comm_enabled = True     # False: skip mirror / flags (TIMING experiments only -- results are wrong)
        self.load(state)

    # ---- setup
    def _uniform_everywhere(self, D):
        """Is the diffusivity ONE constant over the whole tissue?  Every rank enters the same collective exactly once
        (a rank-local shortcut would leave the ranks whose slab holds a scar out of the all_reduce and hang the job):
        global max of [max(D), -min(D)], uniform iff the global max equals the global min."""
        t = torch.stack([D.max(), -D.min()]).to(torch.float64)
        if self.world > 1:
            if t.is_cuda and self.dist.get_backend(self.group) == "gloo":
                t = t.cpu()
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        t = t.cpu()
        return float(t[0]) == -float(t[1])

    def _with_halo(self, own):
        buf = torch.zeros((self.Hb, self.W), dtype=torch.float32, device=self.dev)
        buf[self.o0:self.o1] = own
        if self.world > 1:
            self._setup.exchange([buf])
        return buf

    # ---- state in / out
    def load(self, state):
        """Replace this rank's own rows by ``state`` (v, w, u of shape (H_local, W)) and refresh the neighbours' halos."""
        cur = self.X[self.par]
        for b, x in zip(cur, state):
            b[self.o0:self.o1] = _as_f32(x, self.dev)
        if self.world > 1:
            if self.comm.kind == "peer":
                self.comm.exchange(cur, self.par)
            else:
                self.comm.exchange(cur)
            self.exchanges += 1

    def state(self, copy=True):
        cur = self.X[self.par]
        return State(*[(b[self.o0:self.o1].clone() if copy else b[self.o0:self.o1]) for b in cur])

    def scatter_local(self, state):
        """Accept this rank's own rows (convenience for callers that hold them as a State)."""
        return State(*[_as_f32(x, self.dev) for x in state])

    def gather_local(self, state):
        return state

    def close(self):
        self.comm.close()

    # ---- stepping
    def advance(self, state, t0, t1, copy=True):
        """Euler steps for counter in [t0, t1) on the whole decomposed tissue; returns this rank's rows.

        ``state`` None: continue from the resident state (the result of the previous call); otherwise it is loaded first.
        ``copy`` False returns views into the exchange buffers, valid until the next call that advances."""
        if state is not None:
            self.load(state)
        ints = _is_int(t0) and _is_int(t1)           # the counter's type follows the bounds (stimulus schedule typing)
        t0, t1 = _scalar(t0), _scalar(t1)
        left = int(max(0, np.ceil(t1 - t0)))
        t = int(t0) if ints else t0
        while left > 0:
            steps = []                                   # one group: up to M launches between two halo exchanges
            while left > 0 and len(steps) < self.M:
                steps.append(min(self.T, left))
                left -= steps[-1]
            nxt = 1 - self.par
            if self.comm_enabled:
                self.comm.wait()                         # the halos of X[par] have arrived
            src = self.X[self.par]
            for j, Tj in enumerate(steps):
                last = j == len(steps) - 1
                apron = 4 * sum(steps[j + 1:])           # halo rows that must still be valid after this launch
                dst = self.X[nxt] if last else (self.Y if src is not self.Y else self.X[self.par])
                r0 = 0 if self.top else self.o0 - apron
                r1 = self.Hb if self.bot else self.o1 + apron
                mirror = self.comm.mirror(nxt) if (last and self.world > 1 and self.comm_enabled) else None
                self.be.euler_rows(src, dst, self.D, self.DX, self.DY, self.params, self.stimuli, t, Tj, self.dt, self.dx,
                                   self.top, self.bot, r0, r1, self.uniform, mirror=mirror)
                src = dst
                t += Tj
            if self.world > 1 and self.comm_enabled:
                self.comm.after_last_launch(self.X[nxt], nxt)
                self.exchanges += 1
            self.par = nxt
        return self.state(copy=copy)
