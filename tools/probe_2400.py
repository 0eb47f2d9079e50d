"""2400^2 / 1800^2 / 3000^2 scar-map tissues: streaming-kernel geometry sweep around the planner's choice."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oracle as O
from cardiax_b200 import _lib, options, solve
options.verbose = False


def timed(fn, n=4, warm=1):
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


for H, uniform in ((2400, False), (2400, True), (1800, False), (3000, False)):
    yy, xx = torch.meshgrid(torch.arange(H, device="cuda", dtype=torch.float32), torch.arange(H, device="cuda", dtype=torch.float32), indexing="ij")
    D = torch.full((H, H), 1e-3, device="cuda") if uniform else 1e-4 + 9e-4 * (0.5 + 0.5 * torch.sin(xx / 7.0) * torch.cos(yy / 9.0))
    u = torch.zeros((H, H), device="cuda"); u[100:200, 100:300] = 1.0
    s0 = solve.State(torch.ones((H, H), device="cuda"), torch.ones((H, H), device="cuda"), u)
    best = None
    for nt in (0, 64, 96, 128, 160, 192):
        for rows in ((0,) if nt == 0 else (0, 48, 64, 80, 100, 120, 160, 200)):
            options.cta_threads, options.rows_per_cta, options.kernel = nt, rows, (2 if nt else 0)
            try:
                s = timed(lambda: solve._forward_euler(s0, 0, 100, O.PARAMSETS["5"], D, [], 0.01, 0.01))
            except Exception as e:  # noqa: BLE001
                continue
            g = H * H * 100 / s / 1e9
            p = _lib.last_plan()
            tag = "AUTO" if nt == 0 else ""
            if nt == 0 or best is None or g > best[0]:
                print("%d^2 %s nt=%d rows=%d -> %s: %.1f Gcs/s %s" % (H, "uniform" if uniform else "scar", nt, rows, {k: p[k] for k in ("cta_threads", "strips", "rows_per_cta", "row_chunks", "ctas_per_sm")}, g, tag), flush=True)
            if best is None or g > best[0]:
                best = (g, nt, rows)
    options.cta_threads, options.rows_per_cta, options.kernel = 0, 0, 0
