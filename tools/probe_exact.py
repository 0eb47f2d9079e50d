"""Exact numerics: where does the time go?  Streaming kernel on 4096^2 with different states (development probe)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
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


H = int(os.environ.get("PROBE_H", "4096"))
steps = 20
D = torch.full((H, H), 1e-3, device="cuda")
cases = {}
cases["rest (u = 0, v = w = 1)"] = solve.State(torch.ones((H, H), device="cuda"), torch.ones((H, H), device="cuda"), torch.zeros((H, H), device="cuda"))
g = torch.Generator(device="cuda").manual_seed(0)
cases["random in [0.2, 0.8]"] = solve.State(*[0.2 + 0.6 * torch.rand((H, H), device="cuda", generator=g) for _ in range(3)])
yy, xx = torch.meshgrid(torch.arange(H, device="cuda", dtype=torch.float32), torch.arange(H, device="cuda", dtype=torch.float32), indexing="ij")
sm = 0.5 + 0.4 * torch.sin(xx / 37.0) * torch.cos(yy / 41.0)
cases["smooth in [0.1, 0.9]"] = solve.State(sm.clone(), sm.clone(), sm.clone())
wk = bench.make_fk4096(None, H, H)
cases["bench state at step 0"] = solve.State(*[torch.as_tensor(wk[k]).cuda() for k in "vwu"])
options.numerics = "fast"
cases["bench state after 2000 steps"] = solve._forward_euler(cases["bench state at step 0"], 0, 2000, params.PARAMSET_5, D, [], 0.01, 0.01)
for numerics in ("exact", "fast"):
    options.numerics = numerics
    for name, s0 in cases.items():
        s = timed(lambda: solve._forward_euler(s0, 0, steps, params.PARAMSET_5, D, [], 0.01, 0.01))
        print("%-5s %-32s %s %8.1f Gcell-steps/s" % (numerics, name, _lib.last_kernel(), H * H * steps / s / 1e9))
