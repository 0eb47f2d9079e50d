"""Heun integrator (solve._forward_heun): fused (one tile launch per step) vs unfused (2 rhs + 2 stage kernels)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cardiax_b200 import _lib, options, params, solve

options.verbose = False


def timed(fn, n=3, warm=1):
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


print("FK_HEUN_TILE =", os.environ.get("FK_HEUN_TILE"))
for H, steps in ((256, 200), (512, 200), (1200, 100), (4096, 20)):
    u = torch.zeros((H, H), device="cuda"); u[10:60, 20:90] = 1.0
    s0 = solve.State(torch.ones((H, H), device="cuda"), torch.ones((H, H), device="cuda"), u)
    yy, xx = torch.meshgrid(torch.arange(H, device="cuda", dtype=torch.float32), torch.arange(H, device="cuda", dtype=torch.float32), indexing="ij")
    D = 1e-4 + 9e-4 * (0.5 + 0.5 * torch.sin(xx / 7.0) * torch.cos(yy / 9.0))   # smooth scar-like map (white noise blows up)
    res = {}
    for name, spl, num in (("fast (2 Euler launches + combine)", 0, "fast"), ("exact fused tile", 0, "exact"), ("exact unfused", 1, "exact")):
        options.steps_per_launch, options.numerics = spl, num
        before = _lib.lib().fk_launch_count()
        out = solve._forward_heun(s0, 0, steps, params.PARAMSET_3, D, [], 0.01, 0.01)
        launches = _lib.lib().fk_launch_count() - before
        s = timed(lambda: solve._forward_heun(s0, 0, steps, params.PARAMSET_3, D, [], 0.01, 0.01))
        res[name] = out
        print("heun %d^2 %s: %.1f Gcell-steps/s, %.1f us per step, %d launches per call" % (H, name, H * H * steps / s / 1e9, s / steps * 1e6, launches))
    options.steps_per_launch, options.numerics = 0, "fast"
    print("   exact fused == exact unfused:", all(torch.equal(a, b) for a, b in zip(res["exact fused tile"], res["exact unfused"])),
          " max |fast - exact| = %.2e" % max(float((a - b).abs().max()) for a, b in zip(res["fast (2 Euler launches + combine)"], res["exact fused tile"])))
    s = timed(lambda: solve._forward_euler(s0, 0, steps, params.PARAMSET_3, D, [], 0.01, 0.01))
    print("   euler for comparison: %.1f Gcell-steps/s" % (H * H * steps / s / 1e9))
