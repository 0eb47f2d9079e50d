"""Sweep the resident kernel's tile grid / CTA size on the GPU (development tool).

    python tools/probe_resident.py [fk128 fk256 fk512 fk1024]
One line per (workload, tiles, threads): device microseconds per Euler step and Gcell-steps/s of one n-step call.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import oracle as O  # noqa: E402
from cardiax_b200 import _lib, options, solve, stimulus  # noqa: E402


def run(name, work, kernel, tiles=(0, 0), threads=0, n=2000, numerics="fast", T=0, nc=0, edge=(0, 0)):
    options.verbose = False
    options.numerics, options.kernel, options.steps_per_launch = numerics, kernel, T
    options.cta_threads, options.rows_per_cta, options.tiles, options.cells_per_thread = threads, 0, tiles, nc
    options.edge_tile = edge
    st = solve.State(*[torch.as_tensor(work[k]).cuda() for k in "vwu"])
    D = torch.as_tensor(work["D"]).cuda()
    gs = [stimulus.Stimulus(stimulus.Protocol(*s.protocol), torch.as_tensor(s.field).cuda()) for s in work["stimuli"]]
    P = O.PARAMSETS[work["params"]]
    try:
        s = solve._forward_euler(st, 0, n, P, D, gs, 0.01, 0.01)
        torch.cuda.synchronize()
        best = 1e30
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s2 = solve._forward_euler(s, n, 2 * n, P, D, gs, 0.01, 0.01)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3 / n)
        cells = st.u.numel()
        print("%-7s kernel=%d tiles=%-9s edge=%-9s %-5s %7.2f us/step %7.1f Gcs/s  %s" % (
            name, kernel, tiles, edge, numerics, best, cells / best / 1e3, _lib.last_plan()), flush=True)
        if kernel == 4 and os.environ.get("FK_RES_TIMING"):
            import ctypes
            out = (ctypes.c_ulonglong * 8)()
            _lib.lib().fk_resident_timing(out)
            ns = max(1, out[6])
            print("        cycles/step of CTA 0: ring %.0f | interior %.0f | halo wait+copy %.0f | barrier %.0f | total %.0f"
                  % tuple([out[k] / ns for k in range(4)] + [sum(out[:4]) / ns]), flush=True)
    except Exception as e:  # noqa: BLE001
        print("%-7s kernel=%d tiles=%s thr=%d failed: %s" % (name, kernel, tiles, threads, str(e)[:120]), flush=True)


def main():
    which = [a for a in sys.argv[1:] if not a.startswith("-")] or ["fk128", "fk256", "fk512", "fk1024"]
    short = "--short" in sys.argv
    one = "--one" in sys.argv   # a single run of the planner's choice (for ncu)
    # (tiles, cells per thread, (edge tile rows, edge tile column groups); -1 = even split)
    E = (-1, -1)
    sweeps = {
        "fk128": [((0, 0), 0, (0, 0)), ((8, 16), 2, E), ((8, 8), 2, E), ((12, 12), 2, E), ((16, 8), 1, E), ((8, 8), 4, E)],
        "fk256": [((0, 0), 0, (0, 0)), ((8, 16), 2, E), ((12, 12), 2, E), ((16, 8), 4, E), ((8, 8), 2, E), ((8, 8), 4, E)],
        "fk512": [((0, 0), 0, (0, 0)), ((14, 10), 4, (22, 8)), ((14, 10), 4, (18, 6)), ((12, 12), 4, (20, 5)), ((16, 9), 4, (20, 8)),
                  ((12, 12), 4, (32, 8)), ((14, 10), 4, (0, 0)), ((16, 9), 4, (0, 0))],
        "fk1024": [((0, 0), 0, (0, 0)), ((21, 7), 4, (40, 32)), ((21, 7), 4, (0, 0)), ((16, 9), 4, (0, 0)), ((24, 6), 4, (0, 0)),
                   ((12, 12), 4, (0, 0)), ((18, 8), 4, (0, 0))],
    }
    for name in which:
        if name == "fk128":
            work = bench.make_fk128()
        elif name == "fk512":
            work = bench.make_fk512()
        else:
            n = int(name[2:])
            work = bench.make_fk512()
            from tests import common
            _, D = common.smooth_case((n, n), 0)
            work = dict(v=np.ones((n, n), np.float32), w=np.ones((n, n), np.float32), u=bench.make_fk4096(n, n)["u"], D=D,
                        stimuli=[], params="3")
        if one:
            run(name, work, 4, n=500)
            continue
        for tiles, nc, edge in (sweeps[name][:5] if short else sweeps[name]):
            run(name, work, 4, tiles, 0, nc=nc, edge=edge)
        if not short:
            run(name, work, 3, n=400)
            run(name, work, 2, n=400, T=2)
        run(name, work, 4, numerics="exact", n=500)


if __name__ == "__main__":
    main()
