"""Resident kernel: device time per Euler step and the cycle breakdown of CTA (0, 0) (FK_RES_TIMING=1) -- development tool.
    FK_RES_TIMING=1 python tools/probe_res_timing.py [128 256 512 1024]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from cardiax_b200 import _lib, options, params, solve  # noqa: E402

options.verbose = False
KERNEL = int(os.environ.get("PROBE_KERNEL", "4"))
for n in [int(a) for a in sys.argv[1:]] or [128, 256, 512, 1024]:
    D = torch.as_tensor(bench.scar_map((n, n), 0)).cuda()
    u = torch.zeros((n, n), device="cuda"); u[n // 4:n // 2, n // 4:n // 2] = 1.0
    st = solve.State(torch.ones((n, n), device="cuda"), torch.ones((n, n), device="cuda"), u)
    for numerics in ("fast", "exact"):
        options.numerics, options.kernel = numerics, KERNEL
        steps = 1000
        s = solve._forward_euler(st, 0, steps, params.PARAMSET_3, D, [], 0.01, 0.01)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s = solve._forward_euler(s, steps, 2 * steps, params.PARAMSET_3, D, [], 0.01, 0.01)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / steps
        out = (ctypes.c_ulonglong * 8)()
        _lib.lib().fk_resident_timing(out)
        ns = max(1, out[6])
        print("%4d^2 %-5s %6.2f us/step %6.1f Gcs/s  plan %s" % (n, numerics, us, n * n / us / 1e3, _lib.last_plan()))
        print("        cycles/step of CTA %s: ring %.0f | interior %.0f | halo wait+copy %.0f | barrier %.0f | total %.0f ; slowest warp leaves pass A at %.0f, ring at %.0f, interior at %.0f" % tuple(
            [os.environ.get("FK_RES_TIMING", "0")] + [out[k] / ns for k in range(4)] + [sum(out[:4]) / ns, out[7] / ns, out[4] / ns, out[5] / ns]), flush=True)
