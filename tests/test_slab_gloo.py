"""world_size = 2 and 3 on CPU (gloo): the row-slab halo protocol of cardiax_b200.slab with the EMULATED kernels as the
compute backend.  Checks that the decomposed run is bit-identical to the oracle's run of the whole tissue."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class EmuBackend:
    """cardiax_b200.slab backend that runs fk_euler_rows through tests/emu (CPU)."""
    device = torch.device("cpu")

    def dgrad(self, D, dx, phys_top, phys_bottom):
        from tests.emu import emu
        import ctypes
        Dn = np.ascontiguousarray(D.numpy())
        DX, DY = np.empty_like(Dn), np.empty_like(Dn)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        emu.lib().fk_emu_dgrad(p(Dn), p(DX), p(DY), Dn.shape[0], Dn.shape[1], ctypes.c_float(dx), int(phys_top), int(phys_bottom))
        return torch.from_numpy(DX), torch.from_numpy(DY)

    def euler_rows(self, src, dst, D, DX, DY, params, stimuli, t0, nsteps, dt, dx, phys_top, phys_bottom, row0, row1, uniform,
                   mirror=None):
        assert mirror is None     # the CPU runs exchange with torch.distributed (comm="dist")
        import oracle as O
        from tests.emu import emu
        st = [x.numpy() for x in src]
        stim = [O.Stimulus(O.Protocol(*s.protocol), s.field.numpy()) for s in stimuli]
        got, _ = emu.euler(st, t0, t0 + nsteps, params, D.numpy(), stim, dt, dx, exact=True, T=nsteps, kernel=2, cta_threads=32,
                           phys_top=int(phys_top), phys_bottom=int(phys_bottom), row0=row0, row1=row1, uniform=int(uniform))
        for d, g in zip(dst, got):
            d[row0:row1] = torch.from_numpy(g[row0:row1])


def _worker(rank, world, port, nsteps, M, overlap, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle as O
        from cardiax_b200 import slab
        from cardiax_b200.stimulus import Protocol, Stimulus
        from tests import common
        H, W = 40 * world, 160
        st, D, stim = common.random_case((H, W), seed=12, n_stim=2)
        lo, hi = rank * 40, (rank + 1) * 40
        local = [torch.from_numpy(np.ascontiguousarray(x[lo:hi])) for x in st]
        lstim = [Stimulus(Protocol(*s.protocol), torch.from_numpy(np.ascontiguousarray(s.field[lo:hi]))) for s in stim]
        r = slab.SlabRunner(local, torch.from_numpy(np.ascontiguousarray(D[lo:hi])), O.PARAMSETS["3"], lstim, 0.01, 0.01, rank,
                            world, steps_per_launch=1, halo_launches=M, backend=EmuBackend(), overlap=overlap)
        out = r.advance(None, 0, nsteps)                  # the constructor loaded `local`
        if overlap:
            out = r.advance(None, nsteps, nsteps + 3)     # a second segment continues from the resident state
        else:
            out = r.advance(out, nsteps, nsteps + 3)      # ... or from a state handed back in
        gathered = [None] * world
        dist.all_gather_object(gathered, [x.numpy() for x in out])
        if rank == 0:
            from oracle import c_oracle as C
            ref = C.forward_euler(st, 0, nsteps + 3, O.PARAMSETS["3"], D, stim, 0.01, 0.01)
            ok = all(np.array_equal(np.concatenate([g[k] for g in gathered]), ref[k]) for k in range(3))
            q.put(bool(ok))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,nsteps,M,overlap", [(2, 7, 2, True), (3, 5, 3, True), (2, 4, 1, False)])
def test_slab_decomposition_matches_whole_tissue(world, nsteps, M, overlap):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, _free_port() if r < 0 else PORT[0], nsteps, M, overlap, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


PORT = [0]


@pytest.fixture(autouse=True)
def _port():
    PORT[0] = _free_port()
    yield
