"""Throughput of mid-size tissues just above the resident kernel's range (development tool)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import oracle as O  # noqa: E402
from cardiax_b200 import _lib, options, solve  # noqa: E402
from tests import common  # noqa: E402


def run(n, kernel=0, T=0, steps=400, uniform=False, mg=0, tiles=(0, 0), edge=(0, 0)):
    options.verbose = False
    options.numerics, options.kernel, options.steps_per_launch = "fast", kernel, T
    options.maps_global, options.tiles, options.edge_tile = mg, tiles, edge
    _, D = common.smooth_case((n, n), 0)
    if uniform:
        D = np.full((n, n), 1e-3, np.float32)
    st = solve.State(torch.ones((n, n), device="cuda"), torch.ones((n, n), device="cuda"),
                     torch.as_tensor(bench.make_fk4096(None, n, n)["u"]).cuda())
    D = torch.as_tensor(D).cuda()
    P = O.PARAMSETS["3"]
    try:
        s = solve._forward_euler(st, 0, steps, P, D, [], 0.01, 0.01)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s = solve._forward_euler(s, 0, steps, P, D, [], 0.01, 0.01)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3 / steps)
        print("%5d^2 %s kernel=%d T=%d: %7.2f us/step %7.1f Gcs/s  %s %s" % (
            n, "uniform" if uniform else "hetero ", kernel, T, best, n * n / best / 1e3, _lib.last_kernel(), _lib.last_plan()),
            flush=True)
    except Exception as e:  # noqa: BLE001
        print("%5d^2 kernel=%d T=%d failed: %s" % (n, kernel, T, str(e)[:100]), flush=True)


if __name__ == "__main__":
    run(1024)
    run(1024, mg=1)
    run(1024, mg=1, tiles=(12, 12))
    run(1024, mg=1, tiles=(16, 9))
    for n in (1104, 1200, 1360, 1400):
        run(n)
        run(n, tiles=(12, 12))
        run(n, tiles=(16, 9))
        run(n, tiles=(21, 7))
        run(n, kernel=2, T=1)
