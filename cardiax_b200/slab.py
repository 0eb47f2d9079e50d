"""Row-slab decomposition of ONE large tissue over the GPUs of a node (BASELINE config 5).

Not in the reference (it runs a tissue on a single device); SURVEY 8e defines it.  Rank r owns a
contiguous block of rows; every rank keeps ``halo_launches * 4T`` (default 8 * 8 = 64) halo rows of u, v, w from each
neighbour, so halos are exchanged once per ``halo_launches`` kernel launches (T Euler steps each).
Between exchanges a launch also recomputes the shrinking apron of halo rows it still needs.

Overlap: in the last launch before an exchange the two edge bands (the rows the neighbours are
waiting for) are computed first, their NCCL send/recv (``torch.distributed`` P2P, NVLink) is issued,
and the interior rows are computed while the transfer is in flight.  Physical tissue edges exist
only at the top of rank 0 and at the bottom of the last rank; left/right edges on every rank.

Each launch is ``fk_euler_rows`` (include/fk.h): the same streaming + frame kernels as the
single-GPU path, restricted to a window of output rows, so results are bit-identical to a
single-GPU run of the whole tissue (tests/test_slab_gloo.py checks this on CPU with the emulated
kernels, tests/test_gpu_multi.py on GPUs).
"""
import ctypes

import numpy as np
import torch

from . import _lib, options
from .solve import State, _as_f32, _params_struct, _scalar


class CudaBackend:
    """fk_euler_rows / fk_diffusivity_gradients on the current CUDA device and stream."""

    def __init__(self):
        self.L = _lib.lib()
        self.device = torch.device("cuda", torch.cuda.current_device())
        self._ws = torch.empty(4096, dtype=torch.uint8, device=self.device)

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
                   uniform):
        H, W = D.shape
        P = _params_struct(params)
        arr = (_lib.FkStimulus * max(1, len(stimuli)))()
        for i, s in enumerate(stimuli):
            arr[i] = _lib.FkStimulus(s.field.data_ptr(), _scalar(s.protocol.start), _scalar(s.protocol.duration),
                                     _scalar(s.protocol.period))
        o = _lib.FkOptions()
        self.L.fk_default_options(ctypes.byref(o))
        o.exact = int(options.numerics == "exact")
        o.kernel = int(options.kernel)
        o.cta_threads, o.rows_per_cta = int(options.cta_threads), int(options.rows_per_cta)
        o.phys_top, o.phys_bottom = int(phys_top), int(phys_bottom)
        o.uniform_diffusivity = int(uniform)
        o.safe_division = int(options.safe_division)
        _lib.check(self.L.fk_euler_rows(src[0].data_ptr(), src[1].data_ptr(), src[2].data_ptr(), dst[0].data_ptr(),
                                        dst[1].data_ptr(), dst[2].data_ptr(), D.data_ptr(), DX.data_ptr(), DY.data_ptr(),
                                        H, W, ctypes.byref(P), arr, len(stimuli), float(t0), int(nsteps), np.float32(dt),
                                        np.float32(dx), ctypes.byref(o), int(row0), int(row1), self._ws.data_ptr(),
                                        self._ws.numel(), self._stream()))


class SlabRunner:
    """Advance this rank's rows of a row-decomposed tissue.

    state, diffusivity, stimuli fields: this rank's OWN rows, shape (H_local, W).  All ranks must call
    ``advance`` with the same (t0, t1).  ``backend`` is the compute backend (CUDA by default; the CPU tests inject
    the emulated kernels); ``group`` the torch.distributed process group used for the halo exchange.
    """

    def __init__(self, state, diffusivity, params, stimuli, dt, dx, rank, world, steps_per_launch=0, halo_launches=8,
                 backend=None, group=None, overlap=True):
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
        self.overlap = overlap
        u = _as_f32(state[2], self.dev)
        self.Hl, self.W = u.shape
        if self.Hl < 2 * self.Hh + 8:
            raise ValueError("slab of %d rows is too thin for a halo of %d rows" % (self.Hl, self.Hh))
        self.o0 = 0 if self.top else self.Hh          # first own row inside the buffer
        self.o1 = self.o0 + self.Hl
        self.Hb = self.o1 + (0 if self.bot else self.Hh)
        # static maps with their halos (exchanged once)
        self.D = self._with_halo(_as_f32(diffusivity, self.dev))
        self.uniform = self._uniform_everywhere(self.D)
        self.DX, self.DY = self.be.dgrad(self.D, self.dx, self.top, self.bot)
        self.stimuli = [type(s)(s.protocol, self._with_halo(_as_f32(s.field, self.dev))) for s in stimuli]
        self.buf = [[torch.zeros((self.Hb, self.W), dtype=torch.float32, device=self.dev) for _ in range(3)]
                    for _ in range(2)]
        self.cur = 0

    # ---- communication
    def _uniform_everywhere(self, D):
        """Is the diffusivity ONE constant over the whole tissue?  Every rank enters the same collective exactly once
        (a rank-local shortcut would leave the ranks whose slab holds a scar out of the all_reduce and hang the job):
        global max of [max(D), -min(D)], uniform iff the global max equals the global min."""
        t = torch.stack([D.max(), -D.min()]).to(torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        t = t.cpu()
        return float(t[0]) == -float(t[1])

    def _exchange_ops(self, arrays, rows=None):
        """P2P ops moving `rows` (default: the whole halo depth) own edge rows of each array into the neighbours' halos."""
        d, Hh = self.dist, rows or self.Hh
        ops = []
        for a in arrays:
            if not self.top:
                ops.append(d.P2POp(d.isend, a[self.o0:self.o0 + Hh], self.rank - 1, self.group))
                ops.append(d.P2POp(d.irecv, a[self.o0 - Hh:self.o0], self.rank - 1, self.group))
            if not self.bot:
                ops.append(d.P2POp(d.isend, a[self.o1 - Hh:self.o1], self.rank + 1, self.group))
                ops.append(d.P2POp(d.irecv, a[self.o1:self.o1 + Hh], self.rank + 1, self.group))
        return ops

    def _exchange(self, arrays, wait=True):
        ops = self._exchange_ops(arrays)
        works = self.dist.batch_isend_irecv(ops) if ops else []
        if wait:
            for w in works:
                w.wait()
        return works

    def _with_halo(self, own):
        buf = torch.zeros((self.Hb, self.W), dtype=torch.float32, device=self.dev)
        buf[self.o0:self.o1] = own
        self._exchange([buf])
        return buf

    # ---- stepping
    def scatter_local(self, state):
        """Accept this rank's own rows (convenience for callers that hold them as a State)."""
        return State(*[_as_f32(x, self.dev) for x in state])

    def gather_local(self, state):
        return state

    def advance(self, state, t0, t1):
        """Euler steps for counter in [t0, t1) on the whole decomposed tissue; returns this rank's rows."""
        t0, t1 = _scalar(t0), _scalar(t1)
        n = int(max(0, np.ceil(t1 - t0)))
        src = self.buf[self.cur]
        for b, x in zip(src, state):
            b[self.o0:self.o1] = _as_f32(x, self.dev)
        if n == 0:
            return State(*[b[self.o0:self.o1].clone() for b in src])
        self._exchange(src)
        t, left = t0, n
        while left > 0:
            # one group: up to M launches between two halo exchanges
            steps = []
            while left > 0 and len(steps) < self.M:
                steps.append(min(self.T, left))
                left -= steps[-1]
            for j, Tj in enumerate(steps):
                apron = 4 * sum(steps[j + 1:])           # halo rows that must still be valid after this launch
                src, dst = self.buf[self.cur], self.buf[1 - self.cur]
                r0 = 0 if self.top else self.o0 - apron
                r1 = self.Hb if self.bot else self.o1 + apron
                last = j == len(steps) - 1
                if last and left > 0 and self.overlap and self.world > 1:
                    # edge bands first, their exchange overlaps the interior
                    bands = []
                    if not self.top:
                        bands.append((self.o0, self.o0 + self.Hh))
                    if not self.bot:
                        bands.append((self.o1 - self.Hh, self.o1))
                    for (a, b) in bands:
                        self._rows(src, dst, t, Tj, a, b)
                    works = self._exchange(dst, wait=False)
                    self._rows(src, dst, t, Tj, r0 if self.top else self.o0 + self.Hh,
                               r1 if self.bot else self.o1 - self.Hh)
                    for w in works:
                        w.wait()
                else:
                    self._rows(src, dst, t, Tj, r0, r1)
                    if last and left > 0:
                        self._exchange(dst)
                self.cur = 1 - self.cur
                t += Tj
        out = self.buf[self.cur]
        return State(*[b[self.o0:self.o1].clone() for b in out])

    def _rows(self, src, dst, t, nsteps, row0, row1):
        self.be.euler_rows(src, dst, self.D, self.DX, self.DY, self.params, self.stimuli, t, nsteps, self.dt, self.dx,
                           self.top, self.bot, row0, row1, self.uniform)
