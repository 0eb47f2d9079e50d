"""Times the kernels around the Euler loop on the GPU: snapshot resize, Dormand-Prince attempt, electrogram."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from cardiax_b200 import io, metrics, options, params, solve

options.verbose = False


def timed(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3   # us


for H, out in ((1200, 256), (2400, 512), (600, 128), (4096, 512)):
    st = [torch.rand((H, H), device="cuda") for _ in range(3)]
    us = timed(lambda: io._resize_planes(st, H, H, (out, out)))
    print("resize 3 x %d^2 -> %d^2: %.1f us  (%.0f GB/s of input)" % (H, out, us, 3 * H * H * 4 / us / 1e3))

x = torch.rand((64, 512, 512), device="cuda")
us = timed(lambda: metrics.electrogram(x, (100, 200)))
print("electrogram 64 x 512^2: %.1f us (%.0f GB/s)" % (us, x.numel() * 4 / us / 1e3))

for H in (256, 512, 1024):
    s = solve.init((H, H))
    s = solve.State(s.v, s.w, torch.rand((H, H), device="cuda") * 0.3)
    D = torch.full((H, H), 1e-3, device="cuda")
    options.ode_rtol = options.ode_atol = 1e-5
    torch.cuda.synchronize()
    t0 = time.time()
    o = solve._forward_dormandprince(s, np.array([0.0, 1.0], np.float32), params.PARAMSET_3, D, [], 0.01, 0.01)
    torch.cuda.synchronize()
    dt = time.time() - t0
    st = solve.last_ode_stats
    print("dopri5 %d^2, t in [0, 1], tol 1e-5: %s, %.1f ms, %.1f us per attempt, %.2f Gcell-rhs/s" % (
        H, st, dt * 1e3, dt * 1e6 / max(1, st["attempts"]), st["rhs_evals"] * H * H / dt / 1e9))
