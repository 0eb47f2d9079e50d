"""Multi-process checks of the row-slab decomposition on GPUs: N ranks reproduce the single-GPU run of the whole tissue
bit for bit (both numerics), through the fused peer-memory halo exchange (CUDA IPC + mirror stores of the step kernel)
and through the torch.distributed exchange.

* ``test_slab_peer_exchange_two_processes_one_gpu`` needs ONE GPU: two processes share cuda:0 (gloo is the host-side
  control plane), each maps the other's exchange buffers over CUDA IPC -- the same code path as across NVLink.
* ``test_slab_nccl_matches_single_gpu`` needs >= 2 GPUs (one rank per GPU, NCCL control plane).
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, same_gpu, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dev = 0 if same_gpu else rank
    torch.cuda.set_device(dev)
    if same_gpu:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
    try:
        import oracle as O
        from cardiax_b200 import options, slab, solve
        from cardiax_b200.stimulus import Protocol, Stimulus
        from tests import common
        options.verbose = False
        Hl, W = 96, 640
        st, D = common.smooth_case((Hl * world, W), seed=3)
        _, _, stim = common.random_case((Hl * world, W), seed=3, n_stim=2)
        lo, hi = rank * Hl, (rank + 1) * Hl
        ok, fused = True, 0
        for comm in ("peer", "dist"):
            for numerics in ("exact", "fast"):
                options.numerics = numerics
                local = [torch.as_tensor(np.ascontiguousarray(x[lo:hi])).cuda() for x in st]
                lstim = [Stimulus(Protocol(*s.protocol), torch.as_tensor(np.ascontiguousarray(s.field[lo:hi])).cuda()) for s in stim]
                r = slab.SlabRunner(local, torch.as_tensor(np.ascontiguousarray(D[lo:hi])).cuda(), O.PARAMSETS["3"], lstim, 0.01,
                                    0.01, rank, world, steps_per_launch=2, halo_launches=2, comm=comm)
                out = r.advance(None, 0, 21)               # 5 full groups + one single-launch group of ONE step
                out = r.advance(None, 21, 30)              # continues from the resident state
                gstim = [Stimulus(Protocol(*s.protocol), torch.as_tensor(s.field).cuda()) for s in stim]
                whole = solve.State(*[torch.as_tensor(x).cuda() for x in st])
                ref = solve._forward_euler(whole, 0, 30, O.PARAMSETS["3"], torch.as_tensor(D).cuda(), gstim, 0.01, 0.01)
                ok = ok and all(torch.equal(o, f[lo:hi]) for o, f in zip(out, ref))
                # a state handed back in (reloaded, halos refreshed) and views instead of copies
                out2 = r.advance([f[lo:hi] for f in ref], 30, 37, copy=False)
                ref2 = solve._forward_euler(ref, 30, 37, O.PARAMSETS["3"], torch.as_tensor(D).cuda(), gstim, 0.01, 0.01)
                ok = ok and all(torch.equal(o, f[lo:hi]) for o, f in zip(out2, ref2))
                if comm == "peer":
                    fused += r.be.fused_mirrors
                torch.cuda.synchronize()
                r.close()
        flag = torch.tensor([int(ok), int(fused > 0)])
        if not same_gpu:
            flag = flag.cuda()
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            q.put((bool(flag[0].item()), bool(flag[1].item())))
    finally:
        dist.destroy_process_group()


def _run(world, same_gpu):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, same_gpu, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        if p.is_alive():
            p.terminate()
        assert p.exitcode == 0
    equal, fused = q.get(timeout=10)
    assert equal, "slab result differs from the single-GPU run"
    assert fused, "the streaming kernel never mirrored the edge rows itself (fused exchange not exercised)"


def test_slab_peer_exchange_two_processes_one_gpu():
    """The fused halo exchange (peer-mapped buffers, mirror stores, flags) between two processes on ONE device."""
    _run(2, same_gpu=True)


def test_slab_three_processes_one_gpu():
    """A middle rank has two neighbours (both mirrors in one launch, no physical top / bottom edge)."""
    _run(3, same_gpu=True)


def test_slab_nccl_matches_single_gpu():
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    _run(world, same_gpu=False)
