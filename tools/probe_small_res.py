"""Mailbox resident kernel on small tissues: cells per thread / tile grid sweep (development probe)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
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
for n, grids in ((64, [(0, 0), (8, 8), (8, 4), (4, 4), (16, 8)]), (128, [(0, 0), (16, 8), (8, 8), (16, 4), (12, 12), (16, 9)]),
                 (256, [(0, 0), (16, 8), (12, 12), (16, 9), (8, 16)]), (512, [(0, 0), (12, 12), (16, 9), (9, 16), (18, 8)])):
    shp = (n, n)
    s1 = stimulus.linear(shp, stimulus.Direction.NORTH, 0.2, 20.0, stimulus.Protocol(0, 2, 1e9))
    D = torch.full(shp, 1e-3, device="cuda")
    s0 = solve.init(shp)
    print("%d^2" % n)
    for tiles in grids:
        row = []
        for nc in (1, 2, 4):
            options.kernel, options.tiles, options.cells_per_thread = 4, tiles, nc
            try:
                s = timed(lambda: solve._forward_euler(s0, 0, steps, params.PARAMSET_3, D, [s1], 0.01, 0.01))
                p = _lib.last_plan()
                row.append("nc%d %.2f us (%dx%d of %dx%d, %d thr)" % (nc, s / steps * 1e6, p["tile_rows"], p["tile_cols"], p["tile_h"], p["tile_w"], p["cta_threads"]))
            except Exception as e:
                row.append("nc%d n/a" % nc)
        print("   tiles %-9s " % (tiles,) + " | ".join(row), flush=True)
