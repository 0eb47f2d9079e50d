"""Fixed cost of one solve._forward_euler call on a small tissue (development tool): device and host time of calls of
n Euler steps, for several n; the intercept is the per-call overhead (maps kernel, mailbox memset, launch, Python)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import oracle as O  # noqa: E402
from cardiax_b200 import _lib, options, solve, stimulus  # noqa: E402


def main():
    options.verbose = False
    for name, mk in (("fk512", bench.make_fk512), ("fk128", bench.make_fk128)):
        work = mk()
        st = solve.State(*[torch.as_tensor(work[k]).cuda() for k in "vwu"])
        D = torch.as_tensor(work["D"]).cuda()
        gs = [stimulus.Stimulus(stimulus.Protocol(*s.protocol), torch.as_tensor(s.field).cuda()) for s in work["stimuli"]]
        P = O.PARAMSETS[work["params"]]
        for n in (4, 50, 500):
            s = st
            for _ in range(3):
                s = solve._forward_euler(s, 0, n, P, D, gs, 0.01, 0.01)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            calls = 50
            h0 = time.perf_counter()
            e0.record()
            for i in range(calls):
                s = solve._forward_euler(s, 0, n, P, D, gs, 0.01, 0.01)
            e1.record()
            h1 = time.perf_counter()
            torch.cuda.synchronize()
            print("%s n=%3d: device %.1f us/call, host enqueue %.1f us/call  (%s)" % (
                name, n, e0.elapsed_time(e1) * 1e3 / calls, (h1 - h0) * 1e6 / calls, _lib.last_kernel()), flush=True)


if __name__ == "__main__":
    main()
