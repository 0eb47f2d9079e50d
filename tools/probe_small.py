"""Time the kernels available for small / mid-size tissues on the GPU (development tool).

    python tools/probe_small.py
One line per (workload, kernel, T): device microseconds per Euler step and Gcell-steps/s, no profiling events.
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import oracle as O  # noqa: E402
from cardiax_b200 import _lib, options, solve, stimulus  # noqa: E402


def dev_stim(stimuli):
    if len(stimuli) and isinstance(stimuli[0], list):
        return [dev_stim(s) for s in stimuli]
    return [stimulus.Stimulus(stimulus.Protocol(*s.protocol), torch.as_tensor(s.field).cuda()) for s in stimuli]


def run(name, work, kernel, T, n=400, numerics="fast", uniform=None, nt=0, rh=0):
    options.verbose = False
    options.numerics, options.kernel, options.steps_per_launch = numerics, kernel, T
    options.cta_threads, options.rows_per_cta = nt, rh
    st = solve.State(*[torch.as_tensor(work[k]).cuda() for k in "vwu"])
    D = torch.as_tensor(work["D"]).cuda()
    gs = dev_stim(work["stimuli"])
    P = O.PARAMSETS[work["params"]]
    try:
        s = solve._forward_euler(st, 0, n, P, D, gs, 0.01, 0.01)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0 = time.perf_counter()
        e0.record()
        s = solve._forward_euler(s, n, 2 * n, P, D, gs, 0.01, 0.01)
        e1.record()
        h1 = time.perf_counter()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / n
        cells = st.u.numel()
        print("%-8s kernel=%d T=%d %-5s nt=%3d rh=%3d  %8.2f us/step  %7.1f Gcs/s  host %.2f us/step  %s" % (
            name, kernel, T, numerics, nt, rh, us, cells / us / 1e3, (h1 - h0) * 1e6 / n, _lib.last_plan()), flush=True)
    except Exception as e:  # noqa: BLE001
        print("%-8s kernel=%d T=%d failed: %s" % (name, kernel, T, str(e)[:100]), flush=True)


def main():
    which = sys.argv[1:] or ["fk128", "fk512", "ens256", "fk1024"]
    works = {}
    if "fk128" in which:
        works["fk128"] = bench.make_fk128()
    if "fk512" in which:
        works["fk512"] = bench.make_fk512()
    if "ens256" in which:
        works["ens256"] = bench.make_ens256(128)
    if "fk1024" in which:
        w = bench.make_fk4096(1024, 1024)
        works["fk1024"] = w
    for name, work in works.items():
        for kernel, T in ((0, 0), (3, 0), (2, 1), (2, 2), (2, 3), (1, 1), (1, 2)):
            if kernel == 3 and name == "ens256":
                pass
            run(name, work, kernel, T)
        run(name, work, 0, 0, numerics="exact")


if __name__ == "__main__":
    main()
