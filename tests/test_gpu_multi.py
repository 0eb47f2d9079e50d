"""Multi-GPU (NCCL) check of the row-slab decomposition: N ranks reproduce the single-GPU run of the whole tissue bit
for bit (both numerics).  Needs >= 2 GPUs; skipped otherwise."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import oracle as O
        from cardiax_b200 import options, slab, solve
        from cardiax_b200.stimulus import Protocol, Stimulus
        from tests import common
        options.verbose = False
        Hl, W = 96, 640
        st, D = common.smooth_case((Hl * world, W), seed=3)
        _, _, stim = common.random_case((Hl * world, W), seed=3, n_stim=2)
        ok = True
        for numerics in ("exact", "fast"):
            options.numerics = numerics
            lo, hi = rank * Hl, (rank + 1) * Hl
            local = [torch.as_tensor(np.ascontiguousarray(x[lo:hi])).cuda() for x in st]
            lstim = [Stimulus(Protocol(*s.protocol), torch.as_tensor(np.ascontiguousarray(s.field[lo:hi])).cuda()) for s in stim]
            r = slab.SlabRunner(local, torch.as_tensor(np.ascontiguousarray(D[lo:hi])).cuda(), O.PARAMSETS["3"], lstim, 0.01,
                                0.01, rank, world, steps_per_launch=2, halo_launches=2)
            out = r.advance(local, 0, 21)
            out = r.advance(out, 21, 30)
            gstim = [Stimulus(Protocol(*s.protocol), torch.as_tensor(s.field).cuda()) for s in stim]
            ref = solve._forward_euler(solve.State(*[torch.as_tensor(x).cuda() for x in st]), 0, 30, O.PARAMSETS["3"],
                                       torch.as_tensor(D).cuda(), gstim, 0.01, 0.01)
            ok = ok and all(torch.equal(o, f[lo:hi]) for o, f in zip(out, ref))
        flag = torch.tensor([int(ok)], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            q.put(bool(flag.item()))
    finally:
        dist.destroy_process_group()


def test_slab_nccl_matches_single_gpu():
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True
