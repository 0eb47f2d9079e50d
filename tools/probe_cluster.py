"""Cluster form of the resident kernel against the mailbox form and the streaming kernel (development probe)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from cardiax_b200 import _lib, options, params, solve, stimulus
options.verbose = False


def timed(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e-3


steps = 1000
for n in (64, 128, 192, 256, 320, 384):
    shp = (n, n)
    s1 = stimulus.linear(shp, stimulus.Direction.NORTH, 0.2, 20.0, stimulus.Protocol(0, 2, 1e9))
    D = torch.full(shp, 1e-3, device="cuda")
    s0 = solve.init(shp)
    row = []
    for kernel, tiles, nc in ((5, (0, 0), 0), (5, (4, 4), 2), (5, (4, 4), 4), (5, (2, 4), 0), (5, (2, 2), 0), (4, (0, 0), 0)):
        options.kernel, options.tiles, options.cells_per_thread = kernel, tiles, nc
        try:
            s = timed(lambda: solve._forward_euler(s0, 0, steps, params.PARAMSET_3, D, [s1], 0.01, 0.01))
            p = _lib.last_plan()
            row.append("k%d %s nc%d: %.2f us/step (%dx%d tiles of %dx%d, %d thr, nc %d)" % (
                kernel, tiles, nc, s / steps * 1e6, p["tile_rows"], p["tile_cols"], p["tile_h"], p["tile_w"], p["cta_threads"], p["cells_per_thread"]))
        except Exception as e:
            row.append("k%d %s nc%d: %s" % (kernel, tiles, nc, str(e)[:60]))
    print("%d^2:\n   " % n + "\n   ".join(row))
options.kernel, options.tiles, options.cells_per_thread = 0, (0, 0), 0
# ensembles
for nsim, n in ((128, 256), (64, 128), (9, 256), (18, 256)):
    work = bench.make_ens256(stimulus, nsim) if n == 256 else None
    if work is None:
        shp = (n, n)
        st = solve.State(torch.ones((nsim,) + shp, device="cuda"), torch.ones((nsim,) + shp, device="cuda"), torch.rand((nsim,) + shp, device="cuda") * 0.3)
        D = torch.full((nsim,) + shp, 1e-3, device="cuda") * (1 + 0.1 * torch.rand((nsim,) + shp, device="cuda"))
        stim = []
    else:
        st = solve.State(*[torch.as_tensor(work[k]).cuda() for k in "vwu"])
        D = torch.as_tensor(work["D"]).cuda()
        stim = [[stimulus.Stimulus(stimulus.Protocol(*p), torch.as_tensor(f).cuda()) for p, f in ss] for ss in work["stimuli"]]
    for kernel in (5, 0):
        options.kernel = kernel
        try:
            s = timed(lambda: solve._forward_euler(st, 100, 600, params.PARAMSET_3, D, stim, 0.01, 0.01), n=3, warm=1)
            print("ensemble %d x %d^2 kernel %d: %s %.1f Gcell-steps/s" % (nsim, n, kernel, _lib.last_kernel(), nsim * n * n * 500 / s / 1e9), _lib.last_plan())
        except Exception as e:
            print("ensemble %d x %d^2 kernel %d: %s" % (nsim, n, kernel, str(e)[:100]))
options.kernel = 0
